"""Developer aid: where (in q) the device gamma.ppf leaves 4 ulp of SciPy's.  python tools/gamma_regions.py a"""
import os
import sys

import numpy as np
import scipy.stats as st

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_graph_gpu import ppf_device  # noqa: E402
from probabilit_b200.modeling import OP  # noqa: E402
import gpu_util  # noqa: E402

a = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
q = np.linspace(1e-4, 1 - 1e-4, 200_001)
got, want = ppf_device(OP["PPF_GAMMA"], q, a, 0.0, 1.0), st.gamma(a).ppf(q)
ulp = gpu_util.ulp_diff(got, want)
for lo in np.arange(0, 1, 0.05):
    m = (q >= lo) & (q < lo + 0.05)
    print(f"a={a} q in [{lo:.2f},{lo + 0.05:.2f}): x in [{want[m].min():.3g},{want[m].max():.3g}] frac<=4 {np.mean(ulp[m] <= 4):.3f} "
          f"median {np.median(ulp[m]):.1f} max {ulp[m].max():.0f}")
