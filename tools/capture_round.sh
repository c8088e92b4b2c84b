#!/bin/bash
# One GPU-box call that produces the round's tracked evidence (copied from gpurun_out/ into profiles/ afterwards).
#   tools/capture_round.sh <tag>
set -u
T=${1:-r2m}
O=gpurun_out
python bench.py --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/${T}_bench_reference.json 2>> $O/${T}_bench.err
for c in bmvs tnt synth; do
  python bench.py --config $c --steps 5 --warmup 3 --no-cpu > $O/${T}_bench_$c.json 2>> $O/${T}_bench.err
done
# launch lists: one hot-path step (conditioned features) and one full forward
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/${T}_launches_hot.csv \
    python bench.py --profile-step hot --no-cpu > /dev/null 2>> $O/${T}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/${T}_launches_full.csv \
    python bench.py --profile-step full --no-cpu > /dev/null 2>> $O/${T}_bench.err
# the six W1 launches of the hot-path step, full set (DRAM traffic -> roofline.traffic)
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:warp_corr_h16 -o $O/${T}_w1_step \
    python bench.py --profile-step hot --no-cpu > /dev/null 2>> $O/${T}_bench.err
# the regularisation launches of stage 2 (main net): conv0 pair + 10 layers of the first branch
ncu --set full --clock-control none --import-source on --profile-from-start off -k 'regex:conv_(tc2|kf)' -s 42 -c 11 -o $O/${T}_tc2_stage2 \
    python bench.py --profile-step hot --no-cpu > /dev/null 2>> $O/${T}_bench.err
tail -c 400 $O/${T}_bench.err
ls -la $O/${T}_*
