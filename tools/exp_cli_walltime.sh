#!/bin/bash
# Drop-in comparison at the command line (GPU box): the same Netflix-shaped protobuf files, the same
# flags, wall clock of the whole process (file parse, ingest, epochs, test RMSE per epoch):
#   experimental-mf_b200/mf   --alg mf --iter 15      (this repo, one B200)
#   oracle/_ref/mf_ref        --alg mf --iter 2       (the reference's own main() behind the shims, all host cores)
set -e
D=${TMPDIR:-/tmp}/mfb_cli_$$
mkdir -p $D
./experimental-mf_b200/getdata -w $D/nf --method synth --nu 480189 --nv 17770 --nnz 100000000 --test 0.01 | tail -1
ls -la $D | awk '{print $5, $9}' | tail -2
ARGS="--alg mf --train $D/nf.train --test $D/nf.test --nu 480189 --nv 17770 --dim 128 --eta 2e-2 --lambda 5e-3 --gam 1.0 --bias 2.76"
# MF_STREAM_INGEST=1 (default): the training file is ingested by the first epoch itself (records decoded on the GPU);
# 0: parsed on the host cores and uploaded before the first epoch
for ingest in 1 1 0; do
  s=$(date +%s%N)
  MF_STREAM_INGEST=$ingest MF_TIMING=1 MFB_TIMING=1 ./experimental-mf_b200/mf $ARGS --iter 15 --fly 8 > $D/ours.log 2>&1
  grep "mf_b200:\|load_blocks" $D/ours.log || true
  e=$(date +%s%N)
  echo "mf (B200) 15 epochs, MF_STREAM_INGEST=$ingest: wall $(( (e - s) / 1000000 )) ms; first line: $(grep -m1 "iter#" $D/ours.log); last line: $(tail -1 $D/ours.log)"
done
if [ -z "$SKIP_REF" ]; then
C=$(nproc)
s=$(date +%s%N)
./oracle/_ref/mf_ref $ARGS --iter 2 --fly $C > $D/ref.log 2>&1
e=$(date +%s%N)
echo "mf_ref (reference source + shims, $C cores) 2 epochs: wall $(( (e - s) / 1000000 )) ms; last line: $(tail -1 $D/ref.log)"
fi
rm -rf $D
