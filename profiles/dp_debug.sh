# bounded multi-GPU smoke run of tests/dp_check.py in several configurations (usage: bash profiles/dp_debug.sh <world>)
W=${1:-2}
export DDPM_DEBUG=1 NCCL_DEBUG=WARN
run() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) tests/dp_check.py > gpurun_out/dpdbg_$name.out 2> gpurun_out/dpdbg_$name.err
  echo "$name rc=$?"; tail -1 gpurun_out/dpdbg_$name.out | cut -c1-700; grep -i "error\|trap\|Traceback" gpurun_out/dpdbg_$name.err | head -5
}
run w${W}_p2p_graph_fp32 DP_PREC=fp32 DP_BN_P2P=1 DP_TRAIN_GRAPH=1
run w${W}_p2p_graph_fp16 DP_PREC=fp16 DP_BN_P2P=1 DP_TRAIN_GRAPH=1
run w${W}_ncclbn_fp32 DP_PREC=fp32 DP_BN_P2P=0 DP_TRAIN_GRAPH=1
