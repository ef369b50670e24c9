set -x
mkdir -p gpurun_out
T=${TAG:-c5}
for P in rows masks; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${T}_cfg5_${P}_launches.csv python profiles/one_call.py 5 2 $P > gpurun_out/${T}_cfg5_${P}.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${T}_cfg3_${P}_launches.csv python profiles/one_call.py 3 2 $P > gpurun_out/${T}_cfg3_${P}.log 2>&1
done
