for o in "" "--opt darcy.omega=1.5" "--opt darcy.omega=2.0" "--opt darcy.omega=3.5" "--opt darcy.schur_degree=3" "--opt darcy.schur_degree=4" "--opt darcy.schur_ratio=8" "--opt darcy.schur_ratio=2.5" "--opt darcy.mass_degree=2" "--opt darcy.amg=0"; do
  echo "== $o"; python tools/spe10_scale.py --scale 0.5 --levels 4 --samples 8 --min-level 0 --max-level 1 $o 2>&1 | grep "^level"
done
