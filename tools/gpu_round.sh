#!/bin/bash
# One gpurun call that collects everything a round needs from a B200 box, bounded in time (run from the repository root):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_round.sh full'
#   /usr/local/graft/bin/gpurun --timeout 300 -- 'bash tools/gpu_round.sh quick [pytest -k expression]'
#   /usr/local/graft/bin/gpurun --timeout 300 -- 'bash tools/gpu_round.sh micro backward 46 3'
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/gpu_round.sh variant SDV_BAND_REV'
# Everything lands in gpurun_out/ (merged back by gpurun); copy what should be judged into profiles/.
set -u
mode=${1:-quick}
out=gpurun_out
mkdir -p $out
case "$mode" in
quick)
    timeout 200 python -m pytest tests -m gpu -x -q ${2:+-k "$2"} 2>&1 | tail -15 | tee $out/pytest_gpu.log
    timeout 60 python tools/solve_once.py C3 3 2>&1 | tail -2 | tee $out/solve_once_c3.log
    ;;
full)
    timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $out/pytest_gpu.log
    timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $out/smoke.log
    timeout 300 python bench.py --gpus 1 > $out/bench_c3_n1.json 2> $out/bench_c3_n1.err
    tail -c 600 $out/bench_c3_n1.json
    # launch list of the same command (times under ncu are cold-cache and serialised: shares only, never bench values)
    SDV_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_c3.csv \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c5 > $out/bench_under_ncu.log 2>&1
    ;;
ncu)
    # full capture of one solve's kernels (k_schur, k_chol_band, k_backsub, k_lin_visual): ${2:-C3}
    SDV_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_schur|k_chol_band|k_backsub|k_lin_visual" -c 8 \
        -o $out/ncu_full_${2:-C3} -f python tools/solve_once.py ${2:-C3} > $out/ncu_full.log 2>&1
    tail -3 $out/ncu_full.log
    ;;
variant)
    # build the library with one of the experiment switches of sdv_chol_band.cuh (SDV_BAND_BACKWARD_V2, SDV_BAND_REV,
    # SDV_BAND_BABE) ON THE BOX (the snapshot's default library is not touched in the repository), run the Cholesky-related
    # parity tests and a timed C3 solve:   gpu_round.sh variant SDV_BAND_REV [extra env assignments ...]
    sw=${2:?switch name}
    shift 2
    env "$sw=1" python -c "from sadvio_b200 import build; print(build.build(force=True))" 2>&1 | tail -2
    env "$@" timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "full_solve or band or golden or reference or gradient_tolerance or tight_convergence" 2>&1 | tail -8 | tee $out/variant_$sw.log
    env "$@" timeout 60 python tools/solve_once.py C3 5 2>&1 | tail -1 | tee -a $out/variant_$sw.log
    env "$@" timeout 60 python tools/chol_only.py 2>&1 | tail -3 | tee -a $out/variant_$sw.log
    ;;
traffic)
    # DRAM bytes per launch of every kernel of one solve (bench.py reads profiles/r02_ncu_traffic_*.csv by kernel name)
    SDV_NO_GRAPH=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file $out/ncu_traffic_c3.csv python tools/solve_once.py C3 > $out/ncu_traffic_c3.log 2>&1
    SDV_NO_GRAPH=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        -k regex:"k_lin_visual|k_lin_schur|k_backsub_cost|k_chol_band" -c 12 --log-file $out/ncu_traffic_c5.csv python tools/lin_c5.py C5 3 > $out/ncu_traffic_c5.log 2>&1
    ;;
prebuilt)
    # same checks as `variant` for experiment libraries built in the container (sadvio_b200/_lib/variants/lib_<switch>.so,
    # they travel with the snapshot), selected through SDV_LIB:   gpu_round.sh prebuilt SDV_BAND_REV [env assignments ...]
    sw=${2:?switch name}
    shift 2
    tag=$sw$(echo "$@" | tr -c 'A-Za-z0-9=\n' '_')
    export SDV_LIB=$PWD/sadvio_b200/_lib/variants/lib_$sw.so
    env "$@" timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "full_solve or band or golden or reference or gradient_tolerance or tight_convergence" 2>&1 | tail -8 | tee $out/variant_$tag.log
    env "$@" timeout 60 python tools/solve_once.py C3 5 2>&1 | tail -1 | tee -a $out/variant_$tag.log
    env "$@" timeout 60 python tools/chol_only.py 2>&1 | tail -3 | tee -a $out/variant_$tag.log
    env "$@" timeout 60 python tools/chol_only.py C5 2>&1 | tail -3 | tee -a $out/variant_$tag.log
    ;;
micro)
    name=${2:?micro-benchmark name (tools/micro/<name>.cu)}
    shift 2
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Iinclude -o /tmp/micro_$name tools/micro/$name.cu && timeout 120 /tmp/micro_$name "$@" 2>&1 | tee $out/micro_$name.log
    ;;
*)
    echo "usage: gpu_round.sh quick|full|ncu|variant|micro ..."
    exit 2
    ;;
esac
