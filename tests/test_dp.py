"""CPU, world_size 2 over gloo: the data-parallel plumbing of the policy call.

Each rank takes its row range of a GLOBAL batch, runs the per-rank policy function (here the fp32 oracle stands in for
the CUDA path: the sharding, the shared noise draw and the single all-gather are what is under test) and all-gathers
the action chunks. The gathered result must equal the unsharded call row for row, for even and ragged splits.
"""
from __future__ import annotations

import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, batch: int, out_dir: str):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import vlaser_oracle as O
        from vlaser_b200 import dp
        from vlaser_b200.config import tiny
        from vlaser_b200.synth import make_inputs
        from vlaser_b200.weights import random_state_dict

        cfg = tiny()
        sd = random_state_dict(cfg, 0)
        inp = make_inputs(cfg, batch, 3)
        noise = dp.global_noise(batch, cfg.num_action_tokens, cfg.action_dim, 123)
        seen = {}

        def policy_fn(input_ids, pixel_values, proprios, noise, **_):
            seen["rows"] = input_ids.shape[0]
            assert pixel_values.shape[0] == input_ids.shape[0] * cfg.num_images
            return O.infer_action(cfg, sd, input_ids, pixel_values, proprios, noise)

        glob = {k: inp[k] for k in ("input_ids", "pixel_values", "proprios")}
        got = dp.infer_action_dp(policy_fn, glob, cfg.num_images, noise=noise)
        lo, hi = dp.shard_bounds(batch, world, rank)
        assert seen["rows"] == hi - lo
        assert got.shape == (batch, cfg.horizon_steps, cfg.action_dim)
        # every rank holds the same gathered tensor
        ref0 = got.clone()
        dist.broadcast(ref0, src=0)
        assert torch.equal(ref0, got)
        if rank == 0:
            whole = O.infer_action(cfg, sd, inp["input_ids"], inp["pixel_values"], inp["proprios"], noise)
            torch.save({"got": got, "whole": whole}, os.path.join(out_dir, f"dp_{batch}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [4, 5])
def test_dp_world2_gloo_matches_unsharded(batch, tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), batch, str(tmp_path)), nprocs=world, join=True)
    r = torch.load(os.path.join(str(tmp_path), f"dp_{batch}.pt"))
    # batch rows are independent (test_oracle.py::test_batch_row_independence): only fp32 GEMM blocking differs
    assert float((r["got"] - r["whole"]).abs().max()) <= 2e-5


def test_shard_bounds_cover_and_balance():
    from vlaser_b200.dp import shard_bounds, shard_inputs
    for B in (1, 7, 64, 512):
        for G in (1, 2, 4, 8):
            spans = [shard_bounds(B, G, r) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(G - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    x = {"input_ids": torch.arange(6).view(6, 1), "pixel_values": torch.arange(18).view(18, 1), "scalar": 3}
    s = shard_inputs(x, 3, 2, 1)
    assert s["input_ids"].flatten().tolist() == [3, 4, 5] and s["pixel_values"].flatten().tolist() == list(range(9, 18))
    assert s["scalar"] == 3
