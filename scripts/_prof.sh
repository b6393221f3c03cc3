# bench line, then launch lists + ncu --set full captures of the kernels changed last (each only after its command ran clean)
T=/tmp/prof; mkdir -p $T
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_v6.json 2> gpurun_out/r02_bench_v6.err || { tail -5 gpurun_out/r02_bench_v6.err; exit 1; }
python scripts/train_run.py 6 > gpurun_out/plain_train.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 1400 --csv --log-file gpurun_out/r02c_launch_list_train.csv python scripts/train_run.py 6 > $T/ll.log 2>&1
cap() { name=$1; pat=$2; skip=$3; cnt=$4; shift 4; ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $cnt -o $T/$name "$@" > $T/$name.log 2>&1; python scripts/ncu_summary.py $T/$name.ncu-rep > gpurun_out/r02c_ncu_$name.txt 2>&1; }
cap ce loss_ce_row 2 1 python scripts/train_run.py 4
python scripts/decode_run.py beam 128 2 > gpurun_out/plain_beam.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 0 -c 600 --csv --log-file gpurun_out/r02c_launch_list_beam.csv python scripts/decode_run.py beam 128 2 > $T/llb.log 2>&1
cap beam "beam_|attn_context_mma|attn_head|lstm_beam" 42 7 python scripts/decode_run.py beam 128 2
python scripts/trace_timeline.py beam --concurrent --skip 40 --rows 14 > gpurun_out/r02c_tl_beam.txt 2>&1
python scripts/trace_timeline.py greedy --concurrent --skip 40 --rows 14 > gpurun_out/r02c_tl_greedy.txt 2>&1
python scripts/trace_timeline.py train --rows 60 > gpurun_out/r02c_tl_train.txt 2>&1
ls -la gpurun_out | tail -10
