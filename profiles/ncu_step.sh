set -x
python profiles/profile_step.py > gpurun_out/g1_ps.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/g1_launches.csv python profiles/profile_step.py > /dev/null 2>&1
N=$(python -c "import csv; rows=[r for r in csv.reader(open('gpurun_out/g1_launches.csv')) if len(r)>5 and r[0].isdigit()]; print(len(rows))")
L=$(grep -o "launches/step (ours) [0-9]*" gpurun_out/g1_ps.log | grep -o "[0-9]*$")
SKIP=$((N-L-6))
echo "N=$N L=$L SKIP=$SKIP"
timeout 600 ncu --set full --clock-control none --import-source on --launch-skip $SKIP -c 40 -o /tmp/g1_full python profiles/profile_step.py > gpurun_out/g1_ncu_full.log 2>&1
ncu -i /tmp/g1_full.ncu-rep --page raw --csv > gpurun_out/g1_full_raw.csv
ls -la gpurun_out/ | head
