#!/bin/bash
# final one-GPU evidence run of round 2 (shipped library): GPU suite, smoke, both bench arms, ncu launch list + full capture
OUT=gpurun_out/r2_final2; mkdir -p $OUT
python -m pytest tests -q -m gpu 2>&1 | tail -4 > $OUT/pytest_gpu.txt; cat $OUT/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.txt 2>&1; tail -3 $OUT/smoke.txt
python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
python bench.py --steps 20 --warmup 5 > $OUT/bench_n1_steps20.json 2> $OUT/bench_n1.err; tail -c 200 $OUT/bench_n1_steps20.json
CMD="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vit_decode -s 3 -c 1 -o $OUT/prof_0x011 python scripts/profile_one.py 011 32000000 6 > $OUT/ncu_full.log 2>&1
ls $OUT
