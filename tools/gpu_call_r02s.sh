#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/train_profile.py 8 1 > gpurun_out/r02s_train_profile.txt 2>&1; echo rc=$?
head -80 gpurun_out/r02s_train_profile.txt
