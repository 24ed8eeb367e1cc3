"""TEST INFRASTRUCTURE ONLY — CPU PyTorch port of the reference PMG eval forward, used as the timed
CPU baseline (bench.py `cpu_baseline` leg and `--impl reference`) on the GPU box, where /root/reference
does not exist.  It issues the same ATen ops the reference module issues on CPU (conv2d, batch_norm,
leaky_relu, replication pad, unfold, matmul/div, linear, nn.TransformerEncoder fast path, the 12-task
pooling loop and 12 separate heads) so its wall time is representative of the reference's
(Models/models_PMG.py:176-199); kind = "port" in the bench JSON.

Checked against the golden fixtures in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import warnings
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F


class PMGTorchPort:
    def __init__(self, sd_np: Dict[str, np.ndarray], tasks: Dict[str, int], patch_div: int, nhead: int):
        self.sd = {k: torch.from_numpy(np.array(v)) for k, v in sd_np.items()}
        self.tasks, self.patch_div = dict(tasks), patch_div
        self.convs = []
        i = 0
        while f"feature_extractor.{i}.weight" in self.sd:
            self.convs.append(i)
            i += 3
        d = self.sd["token_proj.bias"].numel()
        nl = 0
        while f"transformer.encoder.layers.{nl}.norm1.weight" in self.sd:
            nl += 1
        self.encoder = None
        if nl:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                layer = torch.nn.TransformerEncoderLayer(d_model=d, nhead=nhead, dim_feedforward=2 * d, dropout=0.1,
                                                         batch_first=True, norm_first=True)   # models_PMG.py:64-68
                self.encoder = torch.nn.TransformerEncoder(layer, num_layers=nl)             # :69
            pre = "transformer.encoder."
            self.encoder.load_state_dict({k[len(pre):]: v for k, v in self.sd.items() if k.startswith(pre)})
            self.encoder.eval()

    @torch.no_grad()
    def __call__(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        sd = self.sd
        for i in self.convs:                                                            # :178
            w = sd[f"feature_extractor.{i}.weight"]
            x = F.conv2d(x, w, stride=w.shape[-1])
            b = f"feature_extractor.{i + 1}"
            x = F.batch_norm(x, sd[f"{b}.running_mean"], sd[f"{b}.running_var"], sd[f"{b}.weight"], sd[f"{b}.bias"],
                             training=False, eps=1e-5)
            x = F.leaky_relu(x, 0.2)
        x = F.conv2d(x, sd["chan_proj.weight"])                                          # :182
        B, C, H, W = x.shape
        ph, pw = (-H) % self.patch_div, (-W) % self.patch_div                            # :157-164
        if ph or pw:
            x = F.pad(x, (0, pw, 0, ph), mode="replicate")
        H, W = x.shape[-2:]
        kh, kw = H // self.patch_div, W // self.patch_div
        p = F.unfold(x, kernel_size=(kh, kw), stride=(kh, kw))                           # :171
        Np = p.shape[-1]
        p = p.view(B, C, kh * kw, Np).permute(0, 3, 1, 2).contiguous()                   # :173
        G = torch.matmul(p, p.transpose(2, 3)) / float(kh * kw)                          # :188
        tok = F.linear(G.view(B, Np, C * C), sd["token_proj.weight"], sd["token_proj.bias"])   # :192-193
        if self.encoder is not None:
            tok = self.encoder(tok)                                                      # :194
        out = {}
        for t in self.tasks:                                                             # :47-57, :198
            q = sd[f"task_pool.query.{t}"].view(1, 1, -1)
            w = (tok * q).sum(-1).softmax(dim=1).unsqueeze(-1)
            out[t] = F.linear((w * tok).sum(1), sd[f"classifiers.{t}.weight"], sd[f"classifiers.{t}.bias"])
        return out
