for v in 0 1 2 3 6 7 8; do echo "== pull_tile variant $v"; IB200_VARIANT=$v python profiles/time_ops.py --ops pull --flags 4 2>&1 | grep -v "^{"; done
