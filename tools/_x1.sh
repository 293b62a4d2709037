mkdir -p gpurun_out/x1
for d in 0 1; do
  QLDPC_X_DISCARD=$d ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:layered_i8s -c 3 --csv --log-file gpurun_out/x1/ncu_d$d.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-fixed10 > gpurun_out/x1/ncu_d$d.log 2>&1
done
