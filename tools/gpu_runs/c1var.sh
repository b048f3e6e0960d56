for f in 16 32 48; do echo "== DRL_C1_FLAGS=$f"; DRL_C1_FLAGS=$f python tools/conv1_timeline.py 2>&1 | grep -A1 "epilogue warp" | tail -1 | cut -c1-200; done
