for cfg in "148 128 0" "148 32 0" "74 32 0" "37 32 0" "148 32 100"; do
set -- $cfg
echo "== grid $1 block $2 sleep $3"
NVNL_ZERO_GRID=$1 NVNL_ZERO_BLOCK=$2 NVNL_ZERO_SLEEP_NS=$3 timeout 120 python profiles/timeline_cfg4.py 2>&1 | grep "prezero=" | tail -3
done
