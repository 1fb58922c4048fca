#!/bin/bash
# whole games with slot refill: 2,048 games on 512 lanes x 50 simulations per move (random-init network)
mkdir -p gpurun_out /tmp/sp_model
timeout 1200 python -m chessrl_b200.selfplay /tmp/sp_model --games ${1:-2048} --lanes ${2:-512} --sims ${3:-50} --no-train > gpurun_out/selfplay_games.log 2>&1
echo "== selfplay: $?"; tail -3 gpurun_out/selfplay_games.log
python - <<PY
import json
d=json.load(open("/tmp/sp_model/gameplays.json"))
import collections
print(len(d), "games; results", collections.Counter(str(g["result"]) for g in d), "mean plies", sum(len(g["moves"]) for g in d)/len(d))
PY
