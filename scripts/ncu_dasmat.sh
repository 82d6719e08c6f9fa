# ncu capture of the translated-plant kernel; the report (with SASS of a ~1 MB kernel) is too large to travel: the raw page is exported on the box
set -e
python scripts/dasmat_check.py --agents 151552 --steps 4 --scenarios trim > gpurun_out/dasmat_plain.log 2>&1
ncu --set full --clock-control none -k regex:dasmat_step_kernel -s 4002 -c 1 -o /tmp/prof_dasmat python scripts/dasmat_check.py --agents 151552 --steps 4 --scenarios trim > gpurun_out/dasmat_ncu.log 2>&1
ncu -i /tmp/prof_dasmat.ncu-rep --page raw --csv > gpurun_out/prof_dasmat_${1:-r02a}_raw.csv
ls -la /tmp/prof_dasmat.ncu-rep >> gpurun_out/dasmat_ncu.log
tail -2 gpurun_out/dasmat_plain.log
