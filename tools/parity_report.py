#!/usr/bin/env python
"""End-to-end parity of the 16-bit product path against the fp32 oracle at BASELINE configs[1] and configs[2], on
asserted-peaky probability volumes (tests/parity_util.py).  Writes one JSON document (default profiles/r02_parity.json):

  python tools/parity_report.py [--out PATH] [--small]

Per config and storage dtype: max / 99.9-percentile / median relative depth error, expected-plane index mismatches (and how
many of those sit within 1e-4 of an integer in the oracle's own sum, hazard H12), confidence error; next to it the same
numbers for (a) the oracle itself with torch's default TF32 convolutions (what the unmodified reference does on this GPU) and
(b) the oracle's fp32 arithmetic with 16-bit storage rounding only (the floor of any kernel that stores in 16 bits)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_parity.json"))
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--peaks", default="0.3", help="comma-separated target peak probabilities")
    args = ap.parse_args()
    import ssmvs_b200
    import parity_util as pu
    ssmvs_b200._lib.bind()
    dev = torch.device("cuda:0")
    doc = {"gpu": torch.cuda.get_device_name(0), "configs": {}}
    v, h, w, d = (5, 128, 160, 48) if args.small else (5, 512, 640, 192)
    for peak in [float(x) for x in args.peaks.split(",")]:
        model, inp, want, cond = pu.peaky_mvsnet(dev, v, h, w, d, seed=0, target_peak=peak)
        row = {"workload": "MVSNet N=%d %dx%d D=%d" % (v, h, w, d), "conditions": cond, "product": {}, "ideal_16bit_storage": {}}
        row["product"]["fp32"] = pu.depth_parity(pu.product_mvsnet(model, inp, torch.float32), want)
        for name, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16)):
            row["product"][name] = pu.depth_parity(pu.product_mvsnet(model, inp, dt), want)
            ideal = pu.ideal_storage_mvsnet(model, inp, dt)
            row["ideal_16bit_storage"][name] = pu.depth_parity(ideal, want)
            row["product"][name + "_vs_ideal_storage"] = pu.depth_parity(pu.product_mvsnet(model, inp, dt), ideal)
        tf, _ = pu.oracle_mvsnet(model, inp, tf32=True)
        row["oracle_tf32_default"] = pu.depth_parity(tf, want)
        doc["configs"]["config2_peak%.2f" % peak] = row
        print(json.dumps({"config2_peak%.2f" % peak: row}, indent=1))
        del model, inp, want
        torch.cuda.empty_cache()
    nsrc, nscale, h, w = (3, 2, 128, 160) if args.small else (4, 3, 512, 640)
    model, inp, want, cond = pu.peaky_cvp(dev, nsrc, nscale, h, w, seed=5, target_peak=0.3)
    row = {"workload": "CVP-MVSNet nsrc=%d nscale=%d %dx%d" % (nsrc, nscale, h, w), "conditions": cond, "product": {}}
    for name, dt in (("fp32", torch.float32), ("fp16", torch.float16), ("bf16", torch.bfloat16)):
        row["product"][name] = pu.cvp_parity(pu.product_cvp(model, inp, dt), want)
    tf, _ = pu.oracle_cvp(model, inp, nscale, tf32=True)
    row["oracle_tf32_default"] = pu.cvp_parity(tf, want)
    doc["configs"]["config3"] = row
    print(json.dumps({"config3": row}, indent=1))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(doc, f, indent=1)


if __name__ == "__main__":
    main()
