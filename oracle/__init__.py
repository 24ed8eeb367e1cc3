"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference PMG / PM forward path.

Nothing in the product package may import this directory.  Allowed importers:
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` (as the checker / the timed CPU baseline, never as the shipped path).
"""
