#!/bin/bash
# timing of every tools/variants/au_*.so (exact assignment, B = 1 and 32, N = 1024) next to the product build
cd $(dirname $0)/..
echo "== product"; python tools/profile_auction.py 2>&1 | cut -d'|' -f1
for so in tools/variants/au_*.so; do echo "== $so"; SHWD_B200_LIB=$PWD/$so python tools/profile_auction.py 2>&1 | cut -d'|' -f1; done
echo "== stage profile"; SHWD_B200_LIB=$PWD/tools/variants/au_prof.so python tools/profile_auction.py 2>&1
