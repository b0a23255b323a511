run() { RENET_DBG_GRU_STEPS=$1 RENET_DBG_GRU_FLAGS=$2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gru_recur --csv python tools/gru_breakdown.py 2>/dev/null | grep gru_recur | tail -1 | awk -F'","' '{print $NF}'; }
echo "steps=0 flags=0: $(run 0 0)"
echo "steps=10 flags=0 (full): $(run 10 0)"
echo "steps=10 no epilogue work (1): $(run 10 1)"
echo "steps=10 no MMA chain (2): $(run 10 2)"
echo "steps=10 no grid barrier (4): $(run 10 4)"
echo "steps=10 no A loads (8): $(run 10 8)"
echo "steps=10 nothing but barrier (11): $(run 10 11)"
