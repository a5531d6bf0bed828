"""oracle/ -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

CPU restatements of the reference hot path of LingmaFuture/PRCV2025REID:

  * oracle.retrieval  -- tools/eval_mm_protocol.py:46-53, 328-365, 389-469, 617-625
  * oracle.sdm        -- models/sdm_loss.py:13-149
  * oracle.ref_loader -- imports the UNMODIFIED reference functions from /root/reference
                         (only possible in the build container; used to pin the restatements
                         and to generate tests/golden/*.npz via oracle/make_golden.py)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package, and there only as the checker / the timed CPU baseline.

Parity pinning: the reference ships NO golden vectors for this path.
The restatements are pinned against outputs of the reference functions themselves, run in the
build container by oracle/make_golden.py and frozen under tests/golden/.
"""
