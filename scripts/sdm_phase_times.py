"""Phase time stamps of tc_fwd_kernel (library built with -DREID_SDM_TIMING): scripts/build_variant.sh WORK x.so -DREID_SDM_TIMING"""
import sys, torch
sys.path.insert(0, '.')
from prcv2025reid_b200 import synth, _cabi
from prcv2025reid_b200 import sdm_loss
P, K = 64, 8
feats, labels = synth.make_sdm_batch(2002, P, K, n_modalities=5, dtype=torch.bfloat16, device="cuda")
y = (labels[:, None] == labels[None, :]).float()
pairs = [(a, b) for a in range(5) for b in range(a)]
qs = [feats[a].clone().requires_grad_(True) for a, b in pairs]
vs = [feats[b].clone().requires_grad_(True) for a, b in pairs]
keep = {}
orig = sdm_loss._SdmPairsFn.forward
def fwd(ctx, *a):
    out = orig(ctx, *a)
    keep["saved"] = ctx.keep[4]
    return out
sdm_loss._SdmPairsFn.forward = staticmethod(fwd)
for it in range(5):
    losses = sdm_loss.sdm_loss_pairs(qs, vs, [y] * 10, tau=0.2)
    losses.sum().backward()
    torch.cuda.synchronize()
L = _cabi.lib()
names = ["prologue", "masks", "accfull", "columns", "rowstats"]
for i in (0, 5, 9):
    sv = keep["saved"][i]
    base = (sv.data_ptr() + 127) // 128 * 128
    off = (base - sv.data_ptr()) // 4
    hdr = off + 8 * 512                                     # den_q..ce_c = 8 arrays of 512 floats
    t = sv[hdr + 80: hdr + 96].view(torch.int64).cpu().tolist()
    print("pair %d fwd:" % i, ", ".join("%s +%.1f us" % (n, (t[k] - t[0]) / 1e3) for k, n in enumerate(names)))
    tb = sv[hdr + 96: hdr + 112].view(torch.int64).cpu().tolist()
    bn = ["prologue", "stats staged", "dS tiles formed", "accfull", "x^ tile loaded", "pass 1", "pass 2"]
    print("pair %d bwd:" % i, ", ".join("%s +%.1f us" % (n, (tb[k] - tb[0]) / 1e3) for k, n in enumerate(bn)))
