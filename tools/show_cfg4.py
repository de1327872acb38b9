"""print the stage times / roofline fractions of bench JSON lines: python tools/show_cfg4.py gpurun_out/TAG_cfg4_s*.json"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = d["stages"]
        print("%s: step %.3f ms | cal %.3f (%.2f) enc %.3f (%.2f) dec %.3f (%.2f) gather %.4f | bits/sym %.3f | value %.3e" % (
            f, d["ms_per_step"], st["calibrate_ms"], st["calibrate_frac"], st["encode_ms"], st["encode_frac"], st["decode_ms"],
            st["decode_frac"], st["gather_ms"], st["bits_per_symbol"], d["value"]))
    except Exception as e:
        print(f, "ERR", e)
