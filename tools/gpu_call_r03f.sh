#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/order_effect.py 2>&1 | tee gpurun_out/r03f_order_effect.txt | tail -20
