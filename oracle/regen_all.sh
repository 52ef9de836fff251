#!/bin/sh
# Regenerate EVERY fixture under tests/golden/ from the UNMODIFIED reference at /root/reference (test infrastructure only).
# Each generator is deterministic (fixed seeds, one torch thread): the regenerated files are bit-identical to the committed ones,
# which is how the oracle's pin can be re-checked:   sh oracle/regen_all.sh && git status --short tests/golden   (no changes)
set -e
cd "$(dirname "$0")/.."
python -m oracle.make_golden buffers dqn sac sac_discrete sac_b256 td3 ddpg bon ppo
python -m oracle.make_golden_dqn_tricks
python -m oracle.make_golden_rainbow
python -m oracle.make_golden_marl maddpg maddpg_bon mappo mappo_disc ippo happo happo_disc
python -m oracle.make_golden_ppo_advance
python -m oracle.make_golden_ppo_tricks
python -m oracle.make_golden_ppo_tricks bon
python -m oracle.make_golden_ppo_tricks beta
python -m oracle.make_golden_siblings
python -m oracle.make_golden_vecloop
python -m oracle.make_golden_mappo_discrete
