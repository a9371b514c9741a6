"""Small calls of the round-2 kernels for compute-sanitizer (memcheck / racecheck), direct launches (no CUDA graph): the
flow-matching loss, greedy generation (causal tcgen05 prefill, split-KV decode attention, argmax), the tcgen05 ViT attention
(257 and 513 tokens), the short and the long tcgen05 prefill attention, the fused tail of the expert projections."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests import ops
from vlaser_b200.config import tiny
from vlaser_b200.policy import PiZeroInference
from vlaser_b200.synth import make_inputs
from vlaser_b200.weights import random_state_dict, text_head_state_dict

DEV = "cuda"
cfg = tiny()
sd = random_state_dict(cfg, 0)
sd.update(text_head_state_dict(cfg, 0))
B = 3
inp = make_inputs(cfg, B, 7, [4, 9, 14])
pol = PiZeroInference(cfg, max_batch=B)
pol.load_state_dict(sd, strict=False)
pol.set_option("use_graph", 0)
g = torch.Generator().manual_seed(3)
act = torch.rand((B, cfg.num_action_tokens, cfg.action_dim), generator=g) * 2 - 1
loss = pol.flow_matching_loss(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), proprios=inp["proprios"].cuda(), actions=act.cuda(),
                              t=torch.rand(B, generator=g).cuda(), noise=inp["noise"].cuda())
print("flow loss", float(loss), flush=True)
for chunk in (128, 64):
    pol.set_option("decode_chunk", chunk)
    toks = pol.generate(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), max_new_tokens=4)
    print("generate", chunk, toks.cpu().tolist(), flush=True)
pol.set_option("fuse_tail", 1)
out = pol.infer_action(inp["input_ids"].cuda(), inp["pixel_values"].cuda(), proprios=inp["proprios"].cuda(), noise=inp["noise"].cuda())
print("infer_action (fuse_tail=1)", float(out.float().abs().max()), flush=True)
del pol

gen = torch.Generator().manual_seed(1)
for N, images, heads in ((257, 2, 2), (513, 2, 2)):
    qkv = torch.randn(images * N, 3 * heads * 64, generator=gen).bfloat16().to(DEV)
    o = ops.vit_attention(qkv, images, N, heads, 0.125, tc=True)
    print("vit_attention_tc", N, float(o.float().abs().max()), flush=True)
nkv, group, hd = 2, 6, 128
nq = nkv * group
for lens, causal, fn in (([85, 130, 256], 0, "short"), ([300, 130, 257], 0, "long"), ([300, 20, 129], 1, "long")):
    cu = [0]
    for n in lens:
        cu.append(cu[-1] + n)
    kvoff = [cu[b] + 5 * b for b in range(len(lens))]
    rows, kv_rows = cu[-1], cu[-1] + 5 * len(lens) + 8
    q = torch.randn(rows, nq * hd, generator=gen).bfloat16().to(DEV)
    kc = torch.randn(nkv, kv_rows, hd, generator=gen).bfloat16().to(DEV)
    vc = torch.randn(nkv, kv_rows, hd, generator=gen).bfloat16().to(DEV)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=DEV)
    if fn == "short":
        o = ops.prefill_attention_tc(q, kc, vc, i32(cu[:-1]), i32(lens), i32(kvoff), nq, nkv, hd ** -0.5)
    else:
        o = ops.prefill_attention_tc_long(q, kc, vc, i32(cu[:-1]), i32(lens), i32(kvoff), nq, nkv, hd ** -0.5, max(lens), 1, causal)
    print("prefill_attention_tc", fn, "causal" if causal else "", float(o.float().abs().max()), flush=True)
torch.cuda.synchronize()
print("done")
