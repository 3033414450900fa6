#!/bin/bash
# A/B of two builds on ONE box (box-to-box variation is +-4 %): personalized_text-to-speech_b200/ab/libvitsdec_{a,b}.so,
# alternating, `bench.py --quick` each.  usage: tools/ab.sh [rounds] [bench args...]
rounds=${1:-3}; shift
for i in $(seq $rounds); do
  for v in a b; do
    VITSDEC_LIB=personalized_text-to-speech_b200/ab/libvitsdec_$v.so python bench.py --quick --steps 20 --warmup 3 "$@" 2>&1 | grep quick | sed "s/^/$v /"
  done
done
