#!/bin/bash
# Round-end measurement suite on one B200 (gpurun): GPU tests, headline bench, ncu launch list, secondary workloads.
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_final_pytest.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_final.csv python tools/profile_step.py --steps 2 > gpurun_out/ncu_final.log 2>&1
python bench.py --mode infer --sweep 1,2,4,8,16,32,64,128,256,512,1024,2048,4096 --no-cpu-baseline > gpurun_out/r02_final_infer.json 2> gpurun_out/r02_final_infer.err
python bench.py --arch crnn --no-cpu-baseline > gpurun_out/r02_final_crnn.json 2> gpurun_out/r02_final_crnn.err
python bench.py --mode stage0 --no-cpu-baseline > gpurun_out/r02_final_stage0_svtr.json 2> gpurun_out/r02_final_stage0_svtr.err
python bench.py --mode stage0 --arch crnn --no-cpu-baseline > gpurun_out/r02_final_stage0_crnn.json 2> gpurun_out/r02_final_stage0_crnn.err
cat gpurun_out/r02_final_pytest.txt
for f in bench infer crnn stage0_svtr stage0_crnn; do python tools/show_bench.py gpurun_out/r02_final_$f.json | head -1; done
# sanitizer passes last: they rebuild the box's copy of the library with the 10-minute mbarrier wait guard
TOOLS="memcheck initcheck" HEAD=6 bash tools/sanitize.sh > gpurun_out/r02_final_sanitizer.log 2>&1
ARCH=crnn B=128 SAN_TIMEOUT=400 timeout 450 compute-sanitizer --tool memcheck --print-limit 10 python tools/sanitize_step.py 2>&1 | grep -v "^$" | tail -6 > gpurun_out/r02_final_sanitizer_crnn.log
tail -3 gpurun_out/r02_final_sanitizer.log; tail -3 gpurun_out/r02_final_sanitizer_crnn.log
