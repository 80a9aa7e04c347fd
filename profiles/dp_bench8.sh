# 8-GPU data-parallel record: correctness at world 4 and 8, then the training step split into compute / exposed
# communication with the timing-only dp_skip option (usage: bash profiles/dp_bench8.sh)
export NCCL_DEBUG=WARN
chk() { # world prec
  DP_PREC=$2 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) tests/dp_check.py > gpurun_out/dp_check_w$1_$2.out 2> gpurun_out/dp_check_w$1_$2.err
  echo "dp_check world=$1 $2 rc=$?"; grep "^{" gpurun_out/dp_check_w$1_$2.out | tail -1 > gpurun_out/dp_check_w$1_$2.json; cut -c1-400 gpurun_out/dp_check_w$1_$2.json
}
chk 8 fp32; chk 8 fp16; chk 4 fp32; chk 4 fp16
for o in "" "dp_skip=3" "dp_skip=1" "dp_skip=2" "bn_p2p=0"; do
  DDPM_OPTS=$o timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29000 + RANDOM % 500)) bench.py --gpus 8 --workload train --steps 50 --warmup 10 --no-cpu > "gpurun_out/r2_dp8_train_$o.json" 2> "gpurun_out/r2_dp8_train_$o.err"
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r2_dp8_train_$o.json')); print('[$o]', round(d['value']), round(d['ms_per_step'],3))
except Exception as e: print('[$o] ERR', open('gpurun_out/r2_dp8_train_$o.err').read()[-800:])
"
done
