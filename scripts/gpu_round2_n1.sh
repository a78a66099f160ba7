#!/bin/bash
# one-GPU evidence run of round 2: full GPU suite, bench (both arms), config5 on one GPU, ncu launch list + full capture
OUT=gpurun_out/r2_final; mkdir -p $OUT
python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > $OUT/pytest_gpu.txt; tail -5 $OUT/pytest_gpu.txt
python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; tail -c 400 $OUT/bench_reference.json
python bench.py --steps 20 --warmup 5 > $OUT/bench_n1_steps20.json 2> $OUT/bench_n1.err; tail -c 300 $OUT/bench_n1_steps20.json
python bench.py --steps 1000 --warmup 5 --no-cpu-baseline > $OUT/bench_n1_steps1000.json 2>> $OUT/bench_n1.err
for w in hard_b32_o32_32M hard_b32_o32_1M s8_f16_o16_256M s16_f16_o16_256M f_b32_o32_4G s8_b16_o32_256M; do
  python bench.py --workload $w --steps 30 --warmup 3 --no-cpu-baseline > $OUT/bench_$w.json 2> $OUT/bench_$w.err
done
python bench.py --impl reference --workload hard_b32_o32_32M --steps 20 --warmup 5 > $OUT/bench_reference_hard_b32.json 2>/dev/null
python bench.py --streams 24 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_s4_24streams.json 2>/dev/null
timeout 300 python bench.py --workload config5 --streams 1024 > $OUT/c5_n1.json 2> $OUT/c5_n1.err; tail -c 300 $OUT/c5_n1.json
# ncu: launch list of the bench command, then the decode kernel in full
CMD="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > $OUT/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
python scripts/profile_one.py 011 32000000 6 > $OUT/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:vit_decode -s 3 -c 1 -o $OUT/prof_0x011 python scripts/profile_one.py 011 32000000 6 > $OUT/ncu_full.log 2>&1
ls -la $OUT | tail -30
