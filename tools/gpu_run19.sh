#!/bin/bash
mkdir -p gpurun_out
MN_MODULE_GRAPHS=0 timeout 300 python tools/profile_conv_layers.py > gpurun_out/r2_conv_layers.txt 2>&1
head -n 50 gpurun_out/r2_conv_layers.txt | cut -c1-200
