"""print a tools/sweep_bench.py result as a table with roofline fractions: python tools/show_sweep.py gpurun_out/TAG_sweep.json [peak GB/s]"""
import json, sys
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6542.1
for r in json.load(open(sys.argv[1])):
    print("BP %3d S %d T %6d | cal %.3f ms (%.2f) | enc %.3f ms (%.2f) | dec %.3f ms (%.2f) | b/sym %.3f | %s" % (
        r["BP"], r["S"], r["T"], r["calibrate_ms"], r["calibrate_GBs"] / peak, r["encode_ms"], r["encode_GBs"] / peak,
        r["decode_ms"], r["decode_GBs"] / peak, r["bits_per_symbol"], "ok" if r["parity_ok"] else "PARITY FAIL"))
