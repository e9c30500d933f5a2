"""Short CPD run for a kernel launch list (ncu): affine 10 iterations + deformable 10 iterations at 5000 x 5000, D = 3."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tools.cpd_bench import problem
from pyfocusr_b200.cpd import affine_registration, deformable_registration

x, y = problem(0, 5000, 5000, 3)
xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
aff = affine_registration(X=xd, Y=yd, max_iterations=10, tolerance=0.0)
aff.register()
reg = deformable_registration(X=xd, Y=aff.transform_point_cloud(yd), max_iterations=10, tolerance=0.0, alpha=0.5, beta=3.0, num_eig=100)
reg.register()
out = reg.transform_point_cloud(torch.from_numpy(problem(1, 15000, 15000, 3)[0]).cuda())
torch.cuda.synchronize()
print("ok", aff.iteration, reg.iteration, float(out.sum()))
