"""Gallery feature store: the on-disk input of the retrieval path.

Reads and writes the reference's cache formats unchanged
    rgb_feats.npy  fp32 [G, 512]  +  rgb_meta.json  list of {"img_id", "pid", "camid"}   (eval_mm_protocol.py:291-325)
    pickle {"g_feat": Tensor[G, D], "g_id": Tensor[G]}                                   (train.py:516-534, 626-631)
and installs them as gallery shards: every rank memory-maps the .npy, copies ONLY its contiguous row range
(`sharding.shard_range`) through a pinned staging buffer to the device in slabs, and runs the one-time
normalise + identity-index pass (`engine.prepare_gallery`).  The 2 GB fp32 file of a 1M-row gallery is therefore
never resident in host memory as a whole, and at N ranks each reads 1/N of it.
"""
import json
import os
import pickle
from typing import List, Tuple

import numpy as np
import torch

from . import engine, sharding

FEATS, META = "rgb_feats.npy", "rgb_meta.json"


def save_cache(cache_dir: str, feats, meta: List[dict]) -> None:
    """Write the reference's cache files (eval_mm_protocol.py:320-323)."""
    os.makedirs(cache_dir, exist_ok=True)
    arr = feats.detach().cpu().numpy() if isinstance(feats, torch.Tensor) else np.asarray(feats)
    np.save(os.path.join(cache_dir, FEATS), arr.astype(np.float32, copy=False))
    with open(os.path.join(cache_dir, META), "w", encoding="utf-8") as f:
        json.dump(meta, f)


def load_meta(cache_dir: str) -> List[dict]:
    with open(os.path.join(cache_dir, META), "r", encoding="utf-8") as f:
        return json.load(f)


def open_feats(cache_dir: str) -> np.ndarray:
    """The feature matrix as a read-only memory map (no copy)."""
    return np.load(os.path.join(cache_dir, FEATS), mmap_mode="r")


def install_shard(feats: np.ndarray, pids, rank: int = 0, world: int = 1, device=None,
                  slab_rows: int = 65536) -> Tuple[engine.GalleryShard, Tuple[int, int]]:
    """Upload rows shard_range(G, rank, world) of `feats` (numpy array or memmap [G, D], any float dtype) in
    pinned slabs and prepare the shard.  `pids`: the WHOLE gallery's person ids [G] (replicated, 8 B per row)."""
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("prcv2025reid_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        device = torch.device("cuda", torch.cuda.current_device())
    G, D = feats.shape
    r0, r1 = sharding.shard_range(G, rank, world)
    dst = torch.empty(r1 - r0, D, dtype=torch.float32, device=device)
    stage = [torch.empty(min(slab_rows, max(1, r1 - r0)), D, dtype=torch.float32).pin_memory() for _ in range(2)]
    done = [None, None]
    for i, s0 in enumerate(range(r0, r1, slab_rows)):
        s1 = min(r1, s0 + slab_rows)
        b = i & 1
        if done[b] is not None:
            done[b].synchronize()                             # the slab's previous H2D copy has drained
        host = stage[b][:s1 - s0]
        host.numpy()[...] = feats[s0:s1]                      # page-in + dtype conversion of this slab only
        dst[s0 - r0:s1 - r0].copy_(host, non_blocking=True)
        done[b] = torch.cuda.Event(); done[b].record()
    pid_t = torch.as_tensor(np.asarray(pids), dtype=torch.int64)
    shard = engine.prepare_gallery(dst, pid_t.to(device), g_offset=r0)
    return shard, (r0, r1)


def load_shard(cache_dir: str, rank: int = 0, world: int = 1, device=None):
    """rgb_feats.npy + rgb_meta.json -> (GalleryShard, meta, (row0, row1))."""
    meta = load_meta(cache_dir)
    feats = open_feats(cache_dir)
    if feats.shape[0] != len(meta):
        raise ValueError("gallery cache is inconsistent: %d feature rows, %d meta entries" % (feats.shape[0], len(meta)))
    shard, rng = install_shard(feats, [int(m["pid"]) for m in meta], rank, world, device)
    return shard, meta, rng


def load_pickle_cache(path: str):
    """train.py:516-534 cache -> (g_feat fp32 numpy [G, D], g_id int64 numpy [G])."""
    with open(path, "rb") as f:
        c = pickle.load(f)
    g_feat, g_id = c["g_feat"], c["g_id"]
    g_feat = g_feat.detach().cpu().numpy() if isinstance(g_feat, torch.Tensor) else np.asarray(g_feat)
    g_id = g_id.detach().cpu().numpy() if isinstance(g_id, torch.Tensor) else np.asarray(g_id)
    return g_feat.astype(np.float32, copy=False), g_id.astype(np.int64, copy=False)
