"""Gallery feature store: the on-disk input of the retrieval path.

Reads and writes the reference's cache formats unchanged
    rgb_feats.npy  fp32 [G, 512]  +  rgb_meta.json  list of {"img_id", "pid", "camid"}   (eval_mm_protocol.py:291-325)
    pickle {"g_feat": Tensor[G, D], "g_id": Tensor[G]}                                   (train.py:516-534, 626-631)
and installs them as gallery shards: every rank memory-maps the .npy, copies ONLY its contiguous row range
(`sharding.shard_range`) through a pinned staging buffer to the device in slabs, and runs the one-time
normalise + identity-index pass (`engine.prepare_gallery`).  The 2 GB fp32 file of a 1M-row gallery is therefore
never resident in host memory as a whole, and at N ranks each reads 1/N of it.

PRE-NORMALISED STORE (`write_store` / `load_store_shard`): the output of that one-time pass kept on disk -- per row range
one raw fp32 file (normalised rows, what the re-scorer reads) and one raw fp16 file (the tensor-core operand copy), the
person ids once, a small JSON index.  Loading a shard is then two memory-mapped files -> pinned slabs -> HBM plus the
identity index: no normalisation kernel, no fp32 -> fp16 pass, no host-side dtype conversion, and 6 B instead of 4 B + a
kernel per element; any rank count can read a store written for another one (row ranges are looked up in the index).
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


# ---------------------------------------------------------------------------------------------
# pre-normalised sharded store
# ---------------------------------------------------------------------------------------------
STORE_INDEX = "store.json"


def write_store(feats, pids, store_dir: str, n_parts: int = 8, device=None, slab_rows: int = 65536) -> dict:
    """Normalise `feats` ([G, D] numpy array / memmap / tensor, un-normalised like the reference's cache) ONCE on the
    device (K1, reid_l2norm_rows) and write the store: part p holds rows shard_range(G, p, n_parts)."""
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("prcv2025reid_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        device = torch.device("cuda", torch.cuda.current_device())
    os.makedirs(store_dir, exist_ok=True)
    G, D = feats.shape
    parts = []
    for p in range(n_parts):
        r0, r1 = sharding.shard_range(G, p, n_parts)
        f32 = np.lib.format.open_memmap(os.path.join(store_dir, "part_%03d.f32.npy" % p), mode="w+", dtype=np.float32, shape=(r1 - r0, D))
        f16 = np.lib.format.open_memmap(os.path.join(store_dir, "part_%03d.f16.npy" % p), mode="w+", dtype=np.float16, shape=(r1 - r0, D))
        for s0 in range(r0, r1, slab_rows):
            s1 = min(r1, s0 + slab_rows)
            x = feats[s0:s1]
            x = x.detach() if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
            n32, n16 = engine.l2norm_rows(x.to(device=device, dtype=torch.float32), want_f16=True)
            f32[s0 - r0:s1 - r0] = n32.cpu().numpy()
            f16[s0 - r0:s1 - r0] = n16.cpu().numpy()
        f32.flush(); f16.flush()
        parts.append([int(r0), int(r1)])
    np.save(os.path.join(store_dir, "pids.npy"), np.asarray(pids, dtype=np.int64))
    index = {"format": "prcv2025reid_b200 pre-normalised gallery store v1", "rows": int(G), "dim": int(D), "parts": parts,
             "normalised": "F.normalize(x, dim=-1) (eval_mm_protocol.py:46-48, :546), fp32 + fp16 copy"}
    with open(os.path.join(store_dir, STORE_INDEX), "w", encoding="utf-8") as f:
        json.dump(index, f)
    return index


def load_store_shard(store_dir: str, rank: int = 0, world: int = 1, device=None, slab_rows: int = 65536):
    """Rows shard_range(G, rank, world) of a pre-normalised store -> (GalleryShard, (row0, row1)).  No arithmetic on the
    features: memory-mapped parts -> pinned slabs -> device, then the identity index (engine.install_normalised)."""
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("prcv2025reid_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        device = torch.device("cuda", torch.cuda.current_device())
    with open(os.path.join(store_dir, STORE_INDEX), "r", encoding="utf-8") as f:
        index = json.load(f)
    G, D = index["rows"], index["dim"]
    r0, r1 = sharding.shard_range(G, rank, world)
    g32 = torch.empty(r1 - r0, D, dtype=torch.float32, device=device)
    g16 = torch.empty(r1 - r0, D, dtype=torch.float16, device=device)
    for kind, dst, dt in (("f32", g32, torch.float32), ("f16", g16, torch.float16)):
        stage = [torch.empty(min(slab_rows, max(1, r1 - r0)), D, dtype=dt).pin_memory() for _ in range(2)]
        done, i = [None, None], 0
        for p, (p0, p1) in enumerate(index["parts"]):
            a, b = max(r0, p0), min(r1, p1)
            if a >= b:
                continue
            part = np.load(os.path.join(store_dir, "part_%03d.%s.npy" % (p, kind)), mmap_mode="r")
            for s0 in range(a, b, slab_rows):
                s1 = min(b, s0 + slab_rows)
                k = i & 1
                if done[k] is not None:
                    done[k].synchronize()
                host = stage[k][:s1 - s0]
                host.numpy()[...] = part[s0 - p0:s1 - p0]
                dst[s0 - r0:s1 - r0].copy_(host, non_blocking=True)
                done[k] = torch.cuda.Event(); done[k].record()
                i += 1
        torch.cuda.current_stream().synchronize()
    pids = torch.from_numpy(np.load(os.path.join(store_dir, "pids.npy")))
    return engine.install_normalised(g32, g16, pids.to(device), g_offset=r0), (r0, r1)
