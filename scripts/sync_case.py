import os, sys, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import exadg_b200
v = int(sys.argv[1])
op = exadg_b200.LaplaceOperator.hypercube(4, 1, 2)
op.set_kernel_variant(v)
x = torch.rand(op.local_size(), dtype=torch.float64, device="cuda")
y = op.initialize_dof_vector()
op.vmult(y, x)
torch.cuda.synchronize()
print("ok variant", v, float(y.abs().sum()))
