T=/tmp/prof; mkdir -p $T
python scripts/train_run.py 6 > gpurun_out/plain_train.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 1400 --csv --log-file gpurun_out/r02_launch_list_train.csv python scripts/train_run.py 6 > $T/ll.log 2>&1
cap() { name=$1; pat=$2; skip=$3; cnt=$4; shift 4; ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $cnt -o $T/$name "$@" > $T/$name.log 2>&1; python scripts/ncu_summary.py $T/$name.ncu-rep > gpurun_out/r02_ncu_$name.txt 2>&1; }
cap ctx attn_context_kernel 60 2 python scripts/train_run.py 5
cap bwd_stream attn_bwd_stream 60 2 python scripts/train_run.py 5
cap dfeat dfeat_gemm 3 1 python scripts/train_run.py 5
cap head attn_head 60 2 python scripts/train_run.py 5
cap gemms tc_gemm_kernel 0 30 python scripts/train_run.py 1
python scripts/decode_run.py beam 128 2 > gpurun_out/plain_beam.log 2>&1
cap beam "beam_|attn_context_mma" 30 6 python scripts/decode_run.py beam 128 2
cp $T/ctx.ncu-rep gpurun_out/r02_prof_ctx.ncu-rep
ls -la $T gpurun_out | tail -30
