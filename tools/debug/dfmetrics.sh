python tools/debug/time_df.py 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:df_ -s 4 -c 2 -o gpurun_out/prof_df -f python tools/debug/time_df.py > /dev/null 2>&1
