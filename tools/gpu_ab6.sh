#!/bin/bash
echo "== chain -3"; MDCTGAN_CHAIN_PRIORITY=-3 timeout 300 python tools/step_time.py 2>&1 | tail -1
echo "== chain -2"; timeout 300 python tools/step_time.py 2>&1 | tail -1
echo "== chain -1"; MDCTGAN_CHAIN_PRIORITY=-1 timeout 300 python tools/step_time.py 2>&1 | tail -1
echo "== chain -3, side -1"; MDCTGAN_CHAIN_PRIORITY=-3 MDCTGAN_SIDE_PRIORITY=-1 timeout 300 python tools/step_time.py 2>&1 | tail -1
echo "== chain -3, side -1, update mid"; MDCTGAN_CHAIN_PRIORITY=-3 MDCTGAN_SIDE_PRIORITY=-1 MDCTGAN_MID_PRIORITY_STREAMS=update,comm timeout 300 python tools/step_time.py 2>&1 | tail -1
