#!/bin/bash
# timing experiments: period of the tcgen05 GDN pipeline with individual stages switched off (results are wrong by design)
for d in 0 1 2 4 8 12 3 6 7 15; do
  echo "== B200VC_GDN_DBG=$d"
  B200VC_GDN_DBG=$d timeout 100 python tools/gdn_time.py 2>&1 | tail -3
done
