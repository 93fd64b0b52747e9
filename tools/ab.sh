#!/bin/bash
# usage: tools/ab.sh out.jsonl variant1 variant2 ... [-- kbench args]   ("main" = the product library)
out=$1; shift
vars=(); while [ $# -gt 0 ] && [ "$1" != "--" ]; do vars+=("$1"); shift; done; [ "$1" == "--" ] && shift
for v in "${vars[@]}"; do
  if [ "$v" == "main" ]; then env -u RRTMGPB_LIB python tools/kbench.py --tag main "$@" >> $out 2>> $out.err
  else RRTMGPB_LIB=rte_rrtmgp_b200/lib/variants/$v.so python tools/kbench.py --tag $v "$@" >> $out 2>> $out.err; fi
done
