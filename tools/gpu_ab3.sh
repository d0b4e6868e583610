#!/bin/bash
echo "== prio on, sweep_D high, side 1"; timeout 300 python tools/step_time.py 2>&1 | tail -1
echo "== prio on, sweep_D low, side 1"; MDCTGAN_LOW_PRIORITY_STREAMS=update,sweep_D timeout 300 python tools/step_time.py 2>&1 | tail -1
echo "== prio on, sweep_D low, side 2"; MDCTGAN_LOW_PRIORITY_STREAMS=update,sweep_D MDCTGAN_SIDE_STREAMS=2 timeout 300 python tools/step_time.py 2>&1 | tail -1
echo "== prio on, sweep_D high, side 2"; MDCTGAN_SIDE_STREAMS=2 timeout 300 python tools/step_time.py 2>&1 | tail -1
echo "== prio on, sweep_D high, side 3"; MDCTGAN_SIDE_STREAMS=3 timeout 300 python tools/step_time.py 2>&1 | tail -1
echo "== prio on, split cap 8"; MDCTGAN_UMMA_MAX_SPLIT=8 timeout 300 python tools/step_time.py 2>&1 | tail -1
