# the two sweeps of substep 0 of frame 40 with every world on its own poses, dataflow form and barrier form
T=${1:-nh}
mkdir -p gpurun_out
for F in 2 0; do
  RP_FLOW=$F timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_solve_(pos|vel)' -c 2 -o gpurun_out/${T}_flow$F -f python bench.py --ncu-frame 40 --hetero > gpurun_out/${T}_flow$F.log 2>&1
  ncu -i gpurun_out/${T}_flow$F.ncu-rep --page raw --csv > gpurun_out/${T}_flow$F.raw.csv 2>/dev/null
done
ls -la gpurun_out/${T}_*
