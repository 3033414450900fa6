#!/bin/bash
# A/B of two option sets of ONE build on ONE box (box-to-box variation is +-4 %), alternating, `bench.py --quick` each.
# usage: tools/ab_opts.sh "<optsA>" "<optsB>" [rounds] [bench args...]      e.g.  tools/ab_opts.sh "mrfp=0" "mrfp=1" 3
A=$1; B=$2; rounds=${3:-3}; shift 3
for i in $(seq $rounds); do
  VITSDEC_OPTS="$A" python bench.py --quick --steps 20 --warmup 3 "$@" 2>&1 | grep quick | sed "s/^/A[$A] /"
  VITSDEC_OPTS="$B" python bench.py --quick --steps 20 --warmup 3 "$@" 2>&1 | grep quick | sed "s/^/B[$B] /"
done
