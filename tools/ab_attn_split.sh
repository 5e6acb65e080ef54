# A/B of the attention tail split (developer tool, run under gpurun): "max parts, min key steps per part"
for cfg in ${AB_CFGS:-"1 2" "0 2" "1 2" "0 2"}; do set -- $cfg
  B200_ATTN_SPLIT=$1 B200_ATTN_SPLIT_MINSTEPS=$2 timeout 200 python bench.py --no-cpu-baseline --steps 30 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('split=$1 minsteps=$2', round(d['value'], 2), round(d['single_sample']['value'], 2), d['clocks']['sm_mhz'])"
done
