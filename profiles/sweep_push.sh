run() { echo "== tile $1 zround $2 persist $3"; IB200_PUSH_TILE=$1 IB200_PUSH_ZROUND=$2 IB200_PUSH_PERSIST=$3 python profiles/time_ops.py --ops push 2>&1 | grep -v "^{"; }
run 432 32 0
run 4324 32 0
run 832 32 0
run 832 4 0
run 816 4 0
run 816 16 0
run 8164 16 0
run 8164 4 0
run 432 32 1
