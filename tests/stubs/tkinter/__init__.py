"""TEST-ONLY stub of `tkinter` (not installed in this image) — see tests/stubs/_pmg_stub_base.py."""
from _pmg_stub_base import Anything, module_getattr

__getattr__ = module_getattr("tkinter")
Tk = Toplevel = Frame = Label = Button = Canvas = Scrollbar = StringVar = IntVar = BooleanVar = Anything
TkVersion = 8.6
END = "end"
