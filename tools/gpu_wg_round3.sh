#!/bin/bash
python tools/unit_error_survey.py 64 0 > gpurun_out/r02_unit_err_wg.txt 2>&1; tail -24 gpurun_out/r02_unit_err_wg.txt
SN_WG=0 python tools/unit_error_survey.py 64 0 > gpurun_out/r02_unit_err_nowg.txt 2>&1; tail -23 gpurun_out/r02_unit_err_nowg.txt
ncu --set full --clock-control none --import-source on -k regex:conv_wg_kernel -c 11 -o gpurun_out/r02_wg_prof -f python tools/wg_profile.py 8 64 > gpurun_out/r02_wg_prof.log 2>&1
tail -3 gpurun_out/r02_wg_prof.log; ls -la gpurun_out/r02_wg_prof.ncu-rep
