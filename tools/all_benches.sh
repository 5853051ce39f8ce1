for prec in tf32x3 fp32; do python bench.py --precision $prec --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_$prec.json; done
python bench.py --farnn 2 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_farnn2.json
python bench.py --mode train --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_train.json
python bench.py --mode train --train-precision fp32 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_train_fp32.json
python tools/bench_configs.py cfg4 cfg5 cfg1 cfg5_onehot > gpurun_out/bench_configs.txt 2>&1
tail -12 gpurun_out/bench_configs.txt
