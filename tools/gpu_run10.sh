#!/bin/bash
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/time_tile.py 3 2>&1 | tail -1
TWXI_KED_CFG=99,99 timeout 300 python tools/time_tile.py 3 2>&1 | tail -1
