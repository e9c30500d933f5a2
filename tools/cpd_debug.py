import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import cpd_port as cp
from test_cpd import _problem
from pyfocusr_b200.cpd import deformable_registration

x, y = _problem(14, 700, 650, 3)
for iters in (1, 2, 5):
    ref = cp.DeformableRegistration(x, y, max_iterations=iters, tolerance=0.0, alpha=0.5, beta=3.0, num_eig=100)
    ref_ty, _ = ref.register()
    reg = deformable_registration(X=x, Y=y, max_iterations=iters, tolerance=0.0, alpha=0.5, beta=3.0, num_eig=100)
    ty, p = reg.register()
    print(iters, "eig", reg.eig_info, "dTY", np.max(np.abs(ty - ref_ty)), "dW", np.max(np.abs(reg.W - ref.W)), np.max(np.abs(ref.W)),
          "sigma2", reg.sigma2, ref.sigma2)
