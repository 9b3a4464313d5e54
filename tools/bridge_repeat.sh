M=tests/golden/models
run() { for i in 1 2 3 4 5 6; do timeout 50 oracle/_ref/ref_bridge $M/$1.xml $2 $3 $4 0.05 10 2>/dev/null | python -c "import sys,json; l=[x for x in sys.stdin if x.startswith('{')]; d=json.loads(l[-1]) if l else {}; print('$1', d.get('rel_linf_vs_reference_cpu'), d.get('worst_variable'), d.get('first_step_over_1e-10'), d.get('bit_identical'))"; done; }
run syn_rules_2d 30 26 1
run syn_rules_3d 14 12 10
run syn_membrane_2d 40 36 1
run syn_membrane_box_2d 30 26 1
run syn_membrane_mixed_2d 40 36 1
run syn_membrane_3d 20 18 16
