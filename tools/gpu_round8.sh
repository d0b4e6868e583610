#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_ddp_gpu.py -x 2>&1 | tail -4
timeout 300 python tools/step_time.py 2>&1 | tail -1
timeout 300 python tools/graph_phases.py 2>&1 | tail -10
