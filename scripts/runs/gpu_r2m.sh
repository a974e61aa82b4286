for c in c3 c2 c5 c4 c1; do scripts/ncu_list.sh $c r2m; done
