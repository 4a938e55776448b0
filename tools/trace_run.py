import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from utmos_b200 import _native, synth
n_vars, n_samples = 1103547, 2504
coh = synth.DeviceCohort(0, n_vars, n_samples)
mask = np.ones(n_samples, np.uint8)
for rep in range(3):
    if rep == 2: os.environ["UTMOS_B200_TRACE"] = "1"
    dm = _native.DeviceMatrix(n_samples, _native.AF_NONE, rows_hint=n_vars)
    dm.append_packed_device(coh.rows.ptr, n_vars, coh.pitch, 0)
    t0 = time.perf_counter(); dm.finalize(); t1 = time.perf_counter()
    dm.begin(mask); t2 = time.perf_counter()
    dm.steps(n_samples); t3 = time.perf_counter()
    dm.close(); t4 = time.perf_counter()
    print("finalize %.3f begin %.3f steps %.3f close %.3f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3), file=sys.stderr)
