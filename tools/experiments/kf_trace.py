"""GPU box: clock64() stamps of CTA 0 of the folded `prob` kernels (kinds PB and PW) at the DTU stage-2 grid: how long does each hop of
the pipeline take?  Events per plane g: 0 producer past `empty`, 1 issuer past `accempty`, 2 issuer past `full` (TMA landed),
3 issuer commits issued, 4 epilogue past `accfull` (the output plane whose centre accumulator is g), 5 its TMEM loads landed,
6 its warp synchronised (release follows)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from dmvsnet_b200 import _native, ops
lib = _native.load()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
d, h, w = 32, 592, 800
prob = ops.PackedLayer((torch.randn(2, 8, 3, 3, 3, generator=g) * 0.1).to(dev), False, None)
x8 = ops.to_ch16(torch.randn(1, 8, d, h, w, device=dev))
trace = torch.zeros(512, 16, dtype=torch.int64, device=dev)
assert lib.dmvs_debug_set_ptr(b"kf_trace", trace.data_ptr()) == 0
NAMES = ["prod past empty", "issuer past accempty", "issuer past full", "issuer committed", "epi past accfull", "epi loads landed", "epi warp synced"]
for label, kf, wide, mw, npr, dbg in (("PB 2 issuers, full", 1, 0, 2, 1, 8), ("PB 2 issuers, skeleton", 1, 0, 2, 1, 11),
                                      ("PW 2 issuers, full", 1, 1, 2, 1, 8), ("PW 2 issuers, skeleton", 1, 1, 2, 1, 11)):
    lib.dmvs_debug_set(b"kf", kf); lib.dmvs_debug_set(b"kf_wide", wide); lib.dmvs_debug_set(b"kf_mw", mw); lib.dmvs_debug_set(b"kf_dbg", dbg)
    for _ in range(3):
        trace.zero_()
        ops.conv3d_ch16(x8, prob, relu=False, out_fmt="f32")
    torch.cuda.synchronize()
    t = trace.cpu().double()
    lo, hi = 64, 448   # steady state
    per_plane = (t[hi, 3] - t[lo, 3]) / (hi - lo)
    R = 10 if wide else 8
    NP = 4
    def m(x):
        return "%6.0f (min %5.0f max %6.0f)" % (x.mean(), x.min(), x.max())
    print("%s: %.0f cycles per plane; issuers %d" % (label, per_plane, mw))
    print("   producer : loop to its next plane %s" % m(t[lo + npr:hi + npr, 8] - t[lo:hi, 7]))
    print("              wait for the stage    %s" % m(t[lo:hi, 0] - t[lo:hi, 8]))
    print("              expect_tx + TMA issue %s" % m(t[lo:hi, 7] - t[lo:hi, 0]))
    print("   TMA      : issued -> issuer past `full` %s" % m(t[lo:hi, 2] - t[lo:hi, 7]))
    print("   issuer   : loop to its next plane %s" % m(t[lo + mw:hi + mw, 9] - t[lo:hi, 3]))
    print("              wait at `accempty`   %s" % m(t[lo:hi, 1] - t[lo:hi, 9]))
    print("              wait at `full`       %s" % m(t[lo:hi, 2] - t[lo:hi, 1]))
    print("              MMAs + commits issued %s" % m(t[lo:hi, 3] - t[lo:hi, 2]))
    print("   epilogue : commit of plane g+1 -> past `accfull` for output g %s" % m(t[lo:hi, 4] - t[lo + 1:hi + 1, 3]))
    print("              previous output's sync -> past `accfull` (stores + waits) %s" % m(t[lo + NP:hi + NP, 4] - t[lo:hi, 6]))
    print("              TMEM loads           %s" % m(t[lo:hi, 5] - t[lo:hi, 4]))
    print("              warp sync            %s" % m(t[lo:hi, 6] - t[lo:hi, 5]))
    print("   slot     : last reader synced (output g+1) -> issuer past `accempty` for plane g+R %s" % m(t[lo + R:hi + R, 1] - t[lo + 1:hi + 1, 6]))
lib.dmvs_debug_set(b"kf", 1); lib.dmvs_debug_set(b"kf_wide", 0); lib.dmvs_debug_set(b"kf_mw", 0); lib.dmvs_debug_set(b"kf_dbg", 0)
lib.dmvs_debug_set_ptr(b"kf_trace", None)
