#!/bin/bash
# A/B of the attention kernel variants inside the real bench step (HY3D_DBG bit 0x100 = separate K/V rings, 0x200 = S released before the exponentials, HY3D_ATTN_POLY = polynomial share)
for cfg in "0 2" "0 3" "0 4" "512 2" "256 2"; do
  set -- $cfg
  HY3D_DBG=$1 HY3D_ATTN_POLY=$2 python bench.py --steps 3 --warmup 3 --cpu-seconds 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('dbg $1 poly $2', round(d['ms_per_step'],2), d['roofline']['families_ms_per_step']['attention'], d['clocks']['sm_mhz'])"
done
