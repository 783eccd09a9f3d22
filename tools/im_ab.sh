# A/B of the two instance-major kernels (run on the GPU box): us/step and % of the measured HBM roofline.
for k in warp cta; do
 for c in ground4 superquadric4 noenv4 ground8 noenv8 superquadric8 ground1 noenv2 superquadric3 ground5 ground12 superquadric32; do
  for n in 65536 1048576; do
   python tools/run_eval.py --case $c --layout instance --im-kernel $k --n $n --steps 200 --warmup 10 --ready
  done
 done
 for c in ground4 ground8; do
  for n in 65536 1048576; do
   python tools/run_eval.py --case $c --layout instance --im-kernel $k --n $n --steps 200 --warmup 10 --perinst
   python tools/run_eval.py --case $c --layout instance --im-kernel $k --n $n --steps 200 --warmup 10 --all
   python tools/run_eval.py --case $c --layout instance --im-kernel $k --n $n --steps 200 --warmup 10 --all --perinst
  done
 done
 python tools/run_eval.py --case ground4 --layout instance --im-kernel $k --n 65536 --steps 200 --warmup 10
done
