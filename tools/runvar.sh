for v in $VARIANTS; do
  G2048_PRINT_LAUNCHES=1 G2048_LIB=$PWD/2048_q-learning_b200/libg2048$v.so python bench.py --steps 20 --warmup 5 --no-extras $BENCH_ARGS > gpurun_out/var$v.json 2> gpurun_out/var$v.err
  python -c "
import json,sys;d=json.load(open('gpurun_out/var$v.json'));t=d['table'];print('variant[$v]',round(d['value']/1e9,2),'G  e2e',round(d['e2e']['value']/1e9,2),'lost',t['lost_update_fraction'],'retried',t.get('retried_update_fraction'),'ms',round(d['roofline']['kernel_ms'],3))"
  grep "per-launch" gpurun_out/var$v.err | cut -c1-200
done
