#!/bin/bash
# Host / PCIe / NUMA topology of the GPU box (explains the e2e ceiling); output under gpurun_out/.
out=gpurun_out/host_probe.txt
{
echo "== lscpu"; lscpu | head -40
echo "== numa nodes"; ls /sys/devices/system/node/ 2>/dev/null
for n in /sys/devices/system/node/node*; do echo "$n: cpus=$(cat $n/cpulist) mem=$(grep MemTotal $n/meminfo)"; done
echo "== nvidia-smi -L"; nvidia-smi -L
echo "== topo"; nvidia-smi topo -m
echo "== gpu numa"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo "$d numa=$(cat $d/numa_node) class=$(cat $d/class) link=$(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null)"; fi; done
echo "== mem"; free -g
echo "== taskset"; taskset -p $$; nproc
echo "== shm"; df -h /dev/shm
echo "== go/javac/node"; which go javac node numactl 2>&1
} > $out 2>&1
