# Split vs whole component-major kernel for the 8-contact shapes (CPLB_CM_KERNEL overrides the dispatch); run on the GPU box.
for k in split whole; do for n in 65536 1048576; do CPLB_CM_KERNEL=$k python tools/run_eval.py --case superquadric8 --layout component --n $n --steps 20 --ready | sed "s/^/$k: /"; done; done
python tools/run_eval.py --case superquadric8 --layout component --n 65536 --steps 20 --ready | sed "s/^/auto: /"
for k in split whole; do for n in 65536 1048576; do CPLB_CM_KERNEL=$k python tools/run_eval.py --case ground8 --layout component --n $n --steps 20 --ready | sed "s/^/$k: /"; done; done
