mkdir -p gpurun_out
bash tools/r02_scale.sh e0 "8"
OIT_EXPERIMENT_SKIP_GEO=1 bash tools/r02_scale.sh e_nogeo "8"
OIT_EXPERIMENT_NO_WAIT=1 bash tools/r02_scale.sh e_nowait "8"
OIT_EXPERIMENT_NO_WAIT=1 OIT_EXPERIMENT_SKIP_GEO=1 bash tools/r02_scale.sh e_nothing "8"
bash tools/r02_scale.sh e_strip16 "8" headline "--strip-rows 16"
bash tools/r02_scale.sh e_strip64 "8" headline "--strip-rows 64"
