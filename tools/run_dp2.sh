#!/bin/bash
# N-GPU data-parallel checks (builder runs, `gpurun --gpus N -- bash tools/run_dp2.sh N [quick]`): the parity tool, then
# scaling lines of the headline workload with global-batch BN statistics exchanged inside the kernels (peer), through
# NCCL, and with local statistics; then the ADMM workload with the feature-sharded global Gram.
N=${1:-2}
TR="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
mkdir -p gpurun_out
$TR tools/dp_parity_check.py > gpurun_out/r02_dp_parity_n${N}.log 2>&1 || tail -30 gpurun_out/r02_dp_parity_n${N}.log
$TR bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r02_dp${N}_resnet20_peer.json 2> gpurun_out/r02_dp${N}_a.err
$TR bench.py --gpus $N --steps 100 --warmup 10 --sync-bn-impl nccl --no-dp-parity > gpurun_out/r02_dp${N}_resnet20_nccl.json 2> gpurun_out/r02_dp${N}_b.err
$TR bench.py --gpus $N --steps 100 --warmup 10 --local-bn > gpurun_out/r02_dp${N}_resnet20_localbn.json 2> gpurun_out/r02_dp${N}_c.err
if [ "$2" != "quick" ]; then
$TR bench.py --gpus $N --steps 20 --warmup 5 --workload resnet56_admm --dp-gram feature --strong > gpurun_out/r02_dp${N}_resnet56_feature_strong.json 2> gpurun_out/r02_dp${N}_d.err
$TR bench.py --gpus $N --steps 20 --warmup 5 --workload resnet56_admm --strong > gpurun_out/r02_dp${N}_resnet56_replica_strong.json 2> gpurun_out/r02_dp${N}_e.err
fi
for f in gpurun_out/r02_dp${N}_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","n_gpus","gpu_launches_per_step")}, d["config"].get("sync_bn"), d["config"].get("sync_bn_impl"))
    print("dp_parity:", d.get("dp_parity"))
except Exception as e:
    print("no json:", e)
PY
done
for f in gpurun_out/r02_dp${N}_?.err; do echo "== $f"; grep -v "^\*\*\*\|OMP_NUM\|^$" $f | tail -8; done
