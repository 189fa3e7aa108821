#!/bin/bash
# usage: tools/try_alt_lib.sh <alt .so or "-"> n_bases seed k   -- runs ksweep with an alternative build of libdnagpu
P=dna-sequences-pg-extension_b200
if [ "$1" != "-" ]; then cp $P/libdnagpu.so /tmp/keep.so; cp "$1" $P/libdnagpu.so; fi
timeout 250 python tools/ksweep.py --n-bases $2 --seed $3 --ks $4 --reps 2 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$1', round(d['ms'],2), round(d['gkmer_s'],2), {k:round(v,2) for k,v in d['kernels'].items()}, d['distinct'])"
if [ "$1" != "-" ]; then cp /tmp/keep.so $P/libdnagpu.so; fi
