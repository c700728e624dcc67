#!/bin/bash
# where the start-up time of readtape_b200 goes on a small capture (B200 box)
xz -dk -c oracle/_ref/examples_full/1600bpi_ukn_6s.tbin.xz > /dev/shm/c.tbin 2>/dev/null || python -c "
import sys; sys.path.insert(0,'.')
from oracle import captures; import shutil; shutil.copy(captures.full_path('1600bpi_ukn_6s'), '/dev/shm/c.tbin')"
nvidia-smi -q | grep -i "persistence mode" | head -1
for i in 1 2 3; do
  /usr/bin/time -f "wall %e s" env RT_STATS=2 RT_TRACE=1 readtape_b200/bin/readtape_b200 -pe -bpi=1600 -ips=50 -tap -q -outf=/dev/shm/o /dev/shm/c.tbin 2>&1 | grep -i "B200 scan\|wall\|\[rt" | head -12
done
cat > /dev/shm/init.cu <<'EOC'
#include <cuda_runtime.h>
#include <stdio.h>
#include <time.h>
static double w(){struct timespec t;clock_gettime(CLOCK_MONOTONIC,&t);return t.tv_sec+1e-9*t.tv_nsec;}
int main(){double t0=w();cudaFree(0);double t1=w();void*p;cudaMalloc(&p,1<<20);double t2=w();void*h;cudaHostAlloc(&h,64<<20,0);double t3=w();
printf("cudaFree(0) %.3f s, cudaMalloc %.3f s, cudaHostAlloc 64MB %.3f s\n",t1-t0,t2-t1,t3-t2);return 0;}
EOC
nvcc -o /dev/shm/init /dev/shm/init.cu 2>/dev/null && for i in 1 2 3; do /dev/shm/init; done
