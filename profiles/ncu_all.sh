# `ncu --set full` captures summarised to CSV on the GPU box (usage: bash profiles/ncu_all.sh sampler|train)
# sampler: every tensor-core / pooling kernel of one U-Net evaluation at the sampler's chunk size (1300 images), the fused
#          last conv + reverse update, the noise kernel;  train: the kernels of one training step at B = 1024.
# Every ncu run is bounded by `timeout`; .ncu-rep files are deleted after the CSV export except the two with source.
O=gpurun_out
T0=$(date +%s)
if [ "$1" = "sampler" ]; then
  python profiles/run_kernel.py forward_infer 1300 1 > $O/plain_fwd.log 2>&1 &&
  timeout 300 ncu --set full --clock-control none -f -k regex:'conv_tc_kernel|conv1f_tc_kernel|bn_apply_pool' -s 48 -c 12 -o $O/ncu_sampler_eval python profiles/run_kernel.py forward_infer 1300 1 > $O/ncu_fwd.log 2>&1
  ncu -i $O/ncu_sampler_eval.ncu-rep --page raw --csv > $O/ncu_sampler_eval_raw.csv 2>/dev/null; rm -f $O/ncu_sampler_eval.ncu-rep
  echo "sampler eval done at $(( $(date +%s) - T0 )) s"
  python profiles/sample_once.py 1300 3 > $O/plain_sample.log 2>&1 &&
  timeout 200 ncu --set full --clock-control none -f -k regex:'conv_tc_kernel<9, 1, 64, 33, 6, 2|randn_dev' -s 2 -c 2 -o $O/ncu_sampler_final python profiles/sample_once.py 1300 3 > $O/ncu_sample.log 2>&1
  ncu -i $O/ncu_sampler_final.ncu-rep --page raw --csv > $O/ncu_sampler_final_raw.csv 2>/dev/null; rm -f $O/ncu_sampler_final.ncu-rep
  echo "sampler final done at $(( $(date +%s) - T0 )) s"
  python profiles/run_kernel.py conv1 1300 3 > $O/plain_c1.log 2>&1 &&
  timeout 120 ncu --set full --clock-control none --import-source on -f -k regex:conv1f_tc_kernel -s 4 -c 1 -o $O/conv1f_tc_r2 python profiles/run_kernel.py conv1 1300 3 > $O/ncu_c1.log 2>&1
  echo "conv1 source done at $(( $(date +%s) - T0 )) s"
else
  python profiles/train_once.py 1024 1 > $O/plain_train.log 2>&1 &&
  # one whole step is ~125 launches over a 3.7 GB working set: `--set full` (40 replays with memory save/restore) took
  # > 15 min for it, so the step is captured with the sections that hold the roofline quantities (~8 replays)
  SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section SchedulerStats"
  MET="--metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum"
  timeout 420 ncu $SEC $MET --clock-control none -f -k regex:'wgrad|bn_bwd_kernel|bn_apply|bn_reduce|l1_bwd|final_bwd|pool_bwd|adam_kernel|qsample|mse_kernel|final_conv|conv1_kernel|conv_tc_kernel|unshuffle|channel_sum|grad_check' -c ${2:-110} -o $O/ncu_train_step python profiles/train_once.py 1024 1 > $O/ncu_train.log 2>&1
  ncu -i $O/ncu_train_step.ncu-rep --page raw --csv > $O/ncu_train_step_raw.csv 2>/dev/null; rm -f $O/ncu_train_step.ncu-rep
  echo "train done at $(( $(date +%s) - T0 )) s"
fi
ls -la $O/ | tail -12
