"""Stage profile of the exact-assignment kernel (needs a -DSHWD_AU_PROFILE variant: tools/build_variant.sh auprof -DSHWD_AU_PROFILE;
run with SHWD_B200_LIB=tools/variants/auprof.so).  The cycle counts of thread 0 land in prices[0..6]."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import shwd, bench
dev = torch.device("cuda:0")
tmpl, src = bench.registration_pairs(32, 1024, 1234, dev)
tmpl = tmpl - tmpl.mean(1, keepdim=True); src = src - src.mean(1, keepdim=True)
for kind in ("sqeuclid", "geodesic"):
    for B in (1, 32):
        for it in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            sig, prices, rounds, status = shwd.ops.exact_assignment(tmpl[:B], src[:B], kind, 2.0, return_info=True)
            torch.cuda.synchronize(); t1 = time.perf_counter()
        p = prices[0, :10].cpu().tolist()
        h = prices[0, 10:40].cpu().tolist()
        tot = sum(p[:4])
        print("%s B=%d: %.2f ms | cycles: start %.0f gs %.0f rescan %.0f apply %.0f (total %.2f ms @1.965GHz) | rescans %d list bids %d rounds %d | %.0f cyc/list bid, %.0f cyc/rescan(16 warps); list evaluations %d, top2 %.0f cyc each, gs %.0f cyc per evaluation, %d of them with a second person waiting" % (
            kind, B, (t1 - t0) * 1e3, p[0], p[1], p[2], p[3], tot / 1.965e6, p[4], p[5], p[6], p[1] / max(p[5], 1), p[2] / max(p[4], 1), p[8], p[7] / max(p[8], 1), p[1] / max(p[8], 1), p[9]))
        if B == 1 and sum(h) > 100:
            names = ["1-8", "9-16", "17-64", "65-511", "512"]
            for q in range(10):
                c, n, pp = h[3 * q:3 * q + 3]
                if n:
                    print("    %s rounds of %s persons: %d rounds, %d rescans, %.2f ms, %.0f cycles per round, %.0f per rescan" % (
                        "list" if q >= 5 else "scan-only", names[q % 5], n, pp, c / 1.965e6, c / n, c / pp))
