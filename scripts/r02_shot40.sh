#!/bin/bash
# direct variant of the host-buffer vmult: parity (child-process test), e2e timings of both variants at several piece sizes; L2-prefetch switch of the k=4 producers
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_host_pipeline.py -q -x > gpurun_out/r02_s40_pytest.log 2>&1; echo "rc $?" >> gpurun_out/r02_s40_pytest.log )
tail -n 4 gpurun_out/r02_s40_pytest.log
cat > /tmp/e2e.py <<'PY'
import sys, time, os, torch
sys.path.insert(0, os.getcwd())
import exadg_b200
torch.cuda.set_device(0)
op = exadg_b200.LaplaceOperator.hypercube(4, 3, 5, 1, 0.0, 2, (0,) * 6, 1.0)
n = op.local_size()
src = torch.rand(n, dtype=torch.float64, device='cuda') * 2 - 1
dst = op.initialize_dof_vector(); op.vmult(dst, src)
h_src = torch.empty(n, dtype=torch.float64).pin_memory(); h_src.copy_(src.cpu())
h_dst = torch.empty(n, dtype=torch.float64).pin_memory()
def timed(fn, reps=5):
    fn(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps
t = timed(lambda: op.vmult_host(h_dst, h_src)); print("sequential %.2f ms %.2f GDoF/s" % (t * 1e3, n / t / 1e9), flush=True)
for mode in sys.argv[1:]:
    op.set_host_pipeline_mode(mode)
    h_dst.fill_(float('nan'))
    t = timed(lambda: op.vmult_host_pipelined(h_dst, h_src))
    ok = (h_dst.cuda() - dst).abs().max().item() == 0.0
    print("%s HS_CELLS=%s HP_CELLS=%s: %.2f ms %.2f GDoF/s bitwise %s" % (mode, os.environ.get("EXADG_B200_HS_CELLS", "-"), os.environ.get("EXADG_B200_HP_CELLS", "-"), t * 1e3, n / t / 1e9, ok), flush=True)
PY
timeout 200 python /tmp/e2e.py staged direct 2>&1 | tail -n 4 | tee gpurun_out/r02_s40_e2e.log
for c in 3072 6144 24576 49152; do
  EXADG_B200_HS_CELLS=$c timeout 120 python /tmp/e2e.py direct 2>&1 | tail -n 1 | tee -a gpurun_out/r02_s40_e2e.log
done
for pf in 0 1; do
  echo "L2PF=$pf" | tee -a gpurun_out/r02_s40_tournament.log
  EXADG_B200_WS_L2PF=$pf TOURNAMENT_VARIANTS=03 timeout 120 build/ws_tournament -1 3 5 100 2>&1 | tail -n 6 | tee -a gpurun_out/r02_s40_tournament.log
done
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv | tee -a gpurun_out/r02_s40_e2e.log
