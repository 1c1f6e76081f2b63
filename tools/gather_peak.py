"""Measured ceiling of random 256-byte row gathers (the propagation kernels' access shape) on this GPU:
table resident in L2 (ML-25M: 56.7 MB) and not (10x config: 563 MB).  Writes profiles/gather_peak.json when
run with --write.  python tools/gather_peak.py [--write]"""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import lgcn_b200  # noqa: E402,F401
from lgcn_b200 import _lib  # noqa: E402


def measure(nrows, ctas, rows_per_hw, reps=7):
    dev = torch.device("cuda:0")
    table = torch.randn(nrows, 64, device=dev)
    sink = torch.zeros(ctas * 16, device=dev)
    L = _lib.lib()
    s = _lib.stream_ptr(dev)

    def run():
        _lib.check(L.lgcn_probe_gather(table.data_ptr(), nrows, ctas, rows_per_hw, sink.data_ptr(), s))
    for _ in range(3):
        run()
    ts = []
    for _ in range(reps):
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); z.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(z))
    ms = float(np.median(ts))
    rows = ctas * 16 * rows_per_hw
    return {"rows_table": nrows, "table_mb": nrows * 256 / 1e6, "ctas": ctas, "rows_gathered": rows, "ms": ms,
            "gather_gbs": rows * 256 / (ms * 1e-3) / 1e9, "rows_per_s": rows / (ms * 1e-3)}


def main():
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    out = {"gpu": torch.cuda.get_device_name(0), "sms": sms, "cases": []}
    for name, nrows in (("l2_resident_ml25m", 221_588), ("hbm_x10", 2_200_000)):
        best = None
        for per_sm in (4, 6, 8):
            r = measure(nrows, sms * per_sm, 4096)
            r["ctas_per_sm"] = per_sm
            if best is None or r["gather_gbs"] > best["gather_gbs"]:
                best = r
        best["name"] = name
        out["cases"].append(best)
        print(name, "%.0f GB/s" % best["gather_gbs"], "at", best["ctas_per_sm"], "CTAs/SM")
    if "--write" in sys.argv:
        with open(os.path.join(REPO, "profiles", "gather_peak.json"), "w") as f:
            json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
