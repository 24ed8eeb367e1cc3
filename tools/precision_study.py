"""PMG precision study (CPU, numpy oracle): which GEMM operands of the fast path can be carried as ONE fp16 plane
(11 significant bits, tcgen05 kind::f16 at the bf16 rate) instead of the bf16 hi/lo pair (~16 bits, 2x the MMAs and bytes).

Rounding points = the tensor-core operands of the CUDA path:
  act0 (conv1's A, smem only)  act1 (HBM, conv2's A)  act2 (chan_proj's operand, smem only)  chan (Gram operand P, smem only)
  gram (G', HBM, token_proj's A)  ln / attn_o / ffh (encoder GEMM operands)  tok_out (heads operand)
Reference = the fp64 oracle on the same bf16-representable inputs.  Usage: python tools/precision_study.py [B] [seeds...]"""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from oracle import synth, pmg_oracle as O

rb = synth.round_bf16


def hilo(a):
    h = rb(a)
    return h + rb(a - h)


def fp16(a):
    return a.astype(np.float16).astype(np.float32)


ROUND = {"bf16": rb, "hilo": hilo, "fp16": fp16, "fp32": lambda v: v}


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))


def run(x, sd, tasks, r):
    f = lambda name, a: ROUND[r.get(name, "hilo")](a.astype(np.float32))
    for (i, kind) in O._plan_from_state_dict(sd):
        w = sd[f"feature_extractor.{i}.weight"]
        x = O.conv_k4s4(x, w) if kind == "strided" else O.conv_1x1(x, w)
        b = f"feature_extractor.{i + 1}"
        x = O.leaky_relu(O.batchnorm_eval(x, sd[f"{b}.weight"], sd[f"{b}.bias"], sd[f"{b}.running_mean"], sd[f"{b}.running_var"]))
        x = f(f"act{i // 3}", x)
    x = f("chan", O.conv_1x1(x, sd["chan_proj.weight"]))
    p, n = O.split_patches(x, 3)
    Gx = O.gram(p, n)
    G = f("gram", Gx)
    B = G.shape[0]
    tok = G.reshape(B, 9, -1) @ sd["token_proj.weight"].T + sd["token_proj.bias"]
    d, nh, hd = 192, 4, 48
    for l in range(2):
        pre = f"transformer.encoder.layers.{l}"
        h = f("ln", O.layer_norm(tok, sd[f"{pre}.norm1.weight"], sd[f"{pre}.norm1.bias"]))
        qkv = h @ sd[f"{pre}.self_attn.in_proj_weight"].T + sd[f"{pre}.self_attn.in_proj_bias"]
        q, k, v = (qkv[..., i * d:(i + 1) * d].reshape(B, 9, nh, hd).transpose(0, 2, 1, 3) for i in range(3))
        att = O.softmax(np.matmul(q, k.transpose(0, 1, 3, 2)) * np.float32(1 / np.sqrt(hd)), -1)
        o = f("attn_o", np.matmul(att, v).transpose(0, 2, 1, 3).reshape(B, 9, d))
        tok = tok + (o @ sd[f"{pre}.self_attn.out_proj.weight"].T + sd[f"{pre}.self_attn.out_proj.bias"])
        h = f("ln", O.layer_norm(tok, sd[f"{pre}.norm2.weight"], sd[f"{pre}.norm2.bias"]))
        ff = f("ffh", np.maximum(h @ sd[f"{pre}.linear1.weight"].T + sd[f"{pre}.linear1.bias"], 0))
        tok = tok + (ff @ sd[f"{pre}.linear2.weight"].T + sd[f"{pre}.linear2.bias"])
    tok = f("tok_out", tok)
    return O.pack_logits(O.task_pool_and_heads(tok, sd, tasks), tasks), Gx


ENC = ("ln", "attn_o", "ffh", "tok_out")
ALL = ("act0", "act1", "act2", "chan", "gram") + ENC


def v(kind, *names):
    return {n: kind for n in names}


variants = {
    "shipped r1 (all hi/lo)": {},
    "all fp16": v("fp16", *ALL),
    "all bf16": v("bf16", *ALL),
    "only act0 fp16": v("fp16", "act0"),
    "only act1 fp16": v("fp16", "act1"),
    "only act2 fp16": v("fp16", "act2"),
    "only chan fp16": v("fp16", "chan"),
    "only gram fp16": v("fp16", "gram"),
    "only encoder fp16": v("fp16", *ENC),
    "backbone (act0/1/2) fp16": v("fp16", "act0", "act1", "act2"),
    "backbone + chan fp16": v("fp16", "act0", "act1", "act2", "chan"),
    "backbone + chan + gram fp16": v("fp16", "act0", "act1", "act2", "chan", "gram"),
    "backbone + gram fp16, chan hilo": v("fp16", "act0", "act1", "act2", "gram"),
    "gram + encoder fp16": v("fp16", "gram", *ENC),
    "act0 + gram fp16": v("fp16", "act0", "gram"),
    "act0 + gram + encoder fp16": v("fp16", "act0", "gram", *ENC),
}

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    seeds = [int(s) for s in sys.argv[2:]] or [0, 5]
    for seed in seeds:
        sd = synth.pmg_state_dict(seed)
        tasks = synth.PMG_TASKS
        x = synth.images(seed, B)
        L64, I64 = O.pmg_forward(x, sd, tasks, 3, 4, dtype=np.float64, return_intermediates=True)
        L64 = O.pack_logits(L64, tasks)
        G64 = I64["gram"]
        print(f"# seed {seed} B {B}: activation ranges  " + "  ".join(f"{k} max|.|={np.abs(I64[k]).max():.3g}" for k in ("act0", "act1", "act2", "chan", "gram", "tokens_in", "tokens_out")))
        for name, r in variants.items():
            L, G = run(x.copy(), sd, tasks, r)
            o = 0
            worst = 0.0
            flips = 0
            for t, n in tasks.items():
                worst = max(worst, rel(L[:, o:o + n], L64[:, o:o + n]))
                flips += int((L[:, o:o + n].argmax(1) != L64[:, o:o + n].argmax(1)).sum())
                o += n
            print(f"seed {seed} {name:36s} logits {rel(L, L64):.2e} worst-task {worst:.2e} gram {rel(G, G64):.2e} argmax flips {flips}/{B * len(tasks)}")
