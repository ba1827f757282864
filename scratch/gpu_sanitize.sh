#!/bin/bash
mkdir -p gpurun_out
for S in 64 25; do
  CC_B200_MLPB=1 PS=$S PB=300 PT=3 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python scratch/prof_mlpb.py > gpurun_out/r02f_memcheck_mlpb_$S.log 2>&1; echo "memcheck mlpb S=$S rc=$?"; tail -2 gpurun_out/r02f_memcheck_mlpb_$S.log
done
CC_B200_MLPB_NH=4 PS=64 PB=300 PT=3 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python scratch/prof_mlpb.py > gpurun_out/r02f_memcheck_mlpb_nh4.log 2>&1; echo "memcheck mlpb nh4 rc=$?"; tail -2 gpurun_out/r02f_memcheck_mlpb_nh4.log
PB=300 PT=40 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python scratch/prof_env.py > gpurun_out/r02f_memcheck_env.log 2>&1; echo "memcheck env rc=$?"; tail -2 gpurun_out/r02f_memcheck_env.log
PB=300 PT=40 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python scratch/prof_env.py > gpurun_out/r02f_racecheck_env.log 2>&1; echo "racecheck env rc=$?"; tail -2 gpurun_out/r02f_racecheck_env.log
PS=64 PB=300 PT=2 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python scratch/prof_mlpb.py > gpurun_out/r02f_racecheck_mlpb.log 2>&1; echo "racecheck mlpb rc=$?"; tail -4 gpurun_out/r02f_racecheck_mlpb.log
PS=64 PB=300 PT=2 timeout 900 compute-sanitizer --tool initcheck --error-exitcode 7 python scratch/prof_mlpb.py > gpurun_out/r02f_initcheck_mlpb.log 2>&1; echo "initcheck mlpb rc=$?"; tail -4 gpurun_out/r02f_initcheck_mlpb.log
