"""Runs simulate_modality (labelled overload) a few times at the full grid; meant to be run under
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:k_sim` for per-kernel times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests._pkg import load
from tests.test_vpa_gpu import phantom

m = load()
img, lab = phantom(160, 192, 160, 1, 1)
for s in range(3):
    out = m.simulate_modality(img[0], lab, 3, 40 + s)
print("ok", float(out.max()))
