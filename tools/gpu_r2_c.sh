#!/bin/bash
# round 2, call C (8 GPUs): which GPUs should carry the host traffic?  memcpy-only combinations
set +e
mkdir -p gpurun_out
P=gpuprefixsums_b200/bin/gps_pcie_probe
{
$P --gpus 4,5 --dir both
$P --gpus 4,5,6 --dir both
$P --gpus 4,5,6,7 --dir h2d
$P --gpus 4,5,6,7 --dir d2h
$P --gpus 0,1,2,3,4,5,6,7 --dirs h2d,h2d,h2d,h2d,d2h,d2h,d2h,d2h
$P --gpus 0,1,2,3,4,5,6,7 --dirs d2h,d2h,d2h,d2h,h2d,h2d,h2d,h2d
$P --gpus 0,1,2,3,4,5,6,7 --dirs h2d,h2d,h2d,h2d,both,both,both,both
$P --gpus 0,1,4,5,6,7 --dirs h2d,h2d,both,both,both,both
$P --gpus 0,4,5,6,7 --dir both
$P --gpus 4,5,6,7 --dir both --mib 64
$P --gpus 4,5,6,7 --dir both --mib 16
$P --gpus 0,1,2,3 --dir both --mib 32
$P --gpus 0,1 --dirs h2d,d2h
$P --gpus 0,1,2,3 --dirs h2d,h2d,d2h,d2h
$P --gpus 4,5,6,7 --dirs h2d,h2d,d2h,d2h
$P --gpus 4,5,6,7 --dir both --drive procs --bind 1
} 2>&1 | tee gpurun_out/r02_pcie_combos.jsonl
lscpu | egrep 'Model name|Socket|NUMA|^CPU\(s\)|Hypervisor' ; nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv,noheader
