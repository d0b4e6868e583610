#!/bin/bash
for e in "A=0" "MDCTGAN_WGRAD_MIN_CTAS=60" "MDCTGAN_WGRAD_MIN_CTAS=60 MDCTGAN_WGRAD_CTA_CAP=74" "MDCTGAN_WGRAD_MIN_CTAS=30 MDCTGAN_WGRAD_CTA_CAP=74" "MDCTGAN_WGRAD_CTA_CAP=100"; do
  echo "== $e"; env $e timeout 300 python tools/step_time.py 2>&1 | tail -1 | cut -c1-60
done
