out=gpurun_out; mkdir -p $out; tag=r01l
WL=synthetic-200taxa-100kpat-1000trees
for rep in 1 2; do for le in 1 0; do
  echo "== BITO_GP_LIBRARY_EXP=$le rep $rep" | tee -a $out/${tag}_ab.log
  BITO_GP_LIBRARY_EXP=$le timeout 300 python tools/time_small.py 2>&1 | grep cuda | tee -a $out/${tag}_ab.log
done; done
for le in 1 0 1 0; do
  echo "== BITO_GP_LIBRARY_EXP=$le" | tee -a $out/${tag}_ab.log
  BITO_GP_LIBRARY_EXP=$le SWEEP_VARIANTS=auto timeout 600 python tools/sweep_variants.py $WL 20000 gauss_seidel 2>&1 | tail -1 | tee -a $out/${tag}_ab.log
done
