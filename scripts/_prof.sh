set -x
python scripts/train_run.py 6 > gpurun_out/plain_train.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 1400 --csv --log-file gpurun_out/r02_launch_list_train.csv python scripts/train_run.py 6 > gpurun_out/ncu_ll.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_context_kernel -s 60 -c 2 -o gpurun_out/r02_prof_ctx python scripts/train_run.py 5 > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_stream -s 60 -c 2 -o gpurun_out/r02_prof_bwd python scripts/train_run.py 5 > gpurun_out/ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dfeat_gemm -s 3 -c 1 -o gpurun_out/r02_prof_dfeat python scripts/train_run.py 5 > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_head -s 60 -c 2 -o gpurun_out/r02_prof_head2 python scripts/train_run.py 5 > gpurun_out/ncu_d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 0 -c 80 -o gpurun_out/r02_prof_gemms python scripts/train_run.py 1 > gpurun_out/ncu_e.log 2>&1
