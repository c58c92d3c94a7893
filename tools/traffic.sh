for e in "A=1" "DECAES_NO_L2_WINDOW=1"; do
env $e ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:voxel_pipeline -c 1 --csv --log-file gpurun_out/traffic_$e.csv python bench.py --voxels 1000000 --steps 1 --warmup 0 --no-e2e --no-cpu > /dev/null 2>&1
echo "$e"; grep -E "dram__|gpu__time" gpurun_out/traffic_$e.csv | awk -F'","' '{print "   ", $(NF-2), $NF}'
done
VOX=400000 bash tools/ab.sh A=1 DECAES_NO_L2_WINDOW=1 2>&1 | head -4
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
