"""Run the FP64 pipe probes (DFMA / DMMA / mixed / latencies) and print JSON."""
import ctypes as C
import json
import sys

sys.path.insert(0, ".")
from jitr_b200 import _cabi

L = _cabi.lib()
out = {}
tf = C.c_double(0)
lat = (C.c_longlong * 8)()
for name, which, w in (("dfma", 0, 0), ("dmma", 1, 0), ("mixed_2of8_dmma", 2, 2), ("mixed_4of8_dmma", 2, 4), ("mixed_6of8_dmma", 2, 6)):
    rc = L.jitr_b200_fp64_probe(which, 4096, w, C.byref(tf), lat)
    out[name + "_tflops"] = tf.value if rc == 0 else None
rc = L.jitr_b200_fp64_probe(3, 1, 0, C.byref(tf), lat)
out["latency_cycles_per_op"] = dict(zip(("dfma", "redux", "shfl64+dadd", "ddiv"), [lat[i] / 1024 for i in range(4)]))
print(json.dumps(out))
