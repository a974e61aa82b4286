WN_LIB=$PWD/wavenets_b200/libwavenet_b200_tl.so python scripts/prof_step.py c3 3 2>&1 | grep -E "SFWD|loss" > gpurun_out/r2j_timeline_c3.log; tail -5 gpurun_out/r2j_timeline_c3.log
WN_LIB=$PWD/wavenets_b200/libwavenet_b200_tl.so python scripts/prof_step.py c2 3 2>&1 | grep -E "SFWD|loss" > gpurun_out/r2j_timeline_c2.log; tail -5 gpurun_out/r2j_timeline_c2.log
