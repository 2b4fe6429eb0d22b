import numpy as np, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import gstools_b200 as gsb
from conftest import synth_modes
cov, z1, z2 = synth_modes(3, 40, seed=1)
axes = [np.arange(8.0), np.arange(16.0), np.arange(130.0)]
gsb.set_option("force_path", 2)
out = gsb.summate_structured(cov, z1, z2, axes)
print(out.sum())
