#!/bin/bash
# development aid: register-liveness dump (nvdisasm -plr) of the TC=10 ADMM kernel -> /tmp/plr.txt
set -e
cd /root/repo/qpcontrol.jl_b200/csrc
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -ccbin /usr/bin/g++ --expt-relaxed-constexpr -DQPC_ONLY_ATLAS "$@" -cubin -o /tmp/only10.cubin api.cu
nvdisasm -plr /tmp/only10.cubin > /tmp/plr_all.txt
awk '/\.text\._Z19qpc_admm_reg_kernelILi${TCSEL:-10}E/{f=1} f' /tmp/plr_all.txt > /tmp/plr.txt
wc -l /tmp/plr.txt
