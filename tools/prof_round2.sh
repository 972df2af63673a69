set -x
N="ncu --set full --clock-control none"
$N --import-source on -k regex:bpr_hogwild -s 3 -c 1 -f -o gpurun_out/prof_hogwild_r2 python tools/prof_targets.py hogwild > gpurun_out/p1.log 2>&1
$N -k regex:bpr_hogwild -s 3 -c 1 -f -o gpurun_out/prof_hogwild_large_r2 python tools/prof_targets.py hogwild_large > gpurun_out/p2.log 2>&1
$N -k regex:score_topk_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_score_c2_r2 python tools/prof_targets.py score_c2 > gpurun_out/p3.log 2>&1
$N --import-source on -k regex:score_topk_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_score_c5_r2 python tools/prof_targets.py score_c5 > gpurun_out/p4.log 2>&1
$N -k regex:gemm_tc_kernel -s 16 -c 8 -f -o gpurun_out/prof_gemm_c3_r2 python tools/prof_targets.py vae > gpurun_out/p5.log 2>&1
$N -k regex:vae_softmax -s 2 -c 1 -f -o gpurun_out/prof_softmax_r2 python tools/prof_targets.py vae > gpurun_out/p6.log 2>&1
$N -k regex:table_reconcile -s 1 -c 1 -f -o gpurun_out/prof_reconcile_r2 python tools/prof_targets.py reconcile > gpurun_out/p7.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --blocks scoring > gpurun_out/p8.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vae_r2.csv python tools/prof_targets.py vae > gpurun_out/p9.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_neumf_r2.csv python tools/prof_targets.py neumf > gpurun_out/p10.log 2>&1
tail -2 gpurun_out/p*.log
ls -la gpurun_out/*.ncu-rep
