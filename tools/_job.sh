O=gpurun_out/r3k; mkdir -p $O
python tools/scrfd_pass.py > $O/scrfd_plain.log 2>&1 && ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc2 -s 3 -c 2 -o $O/scrfd_128 python tools/scrfd_pass.py > $O/ncu_scrfd_full.log 2>&1
tail -2 $O/scrfd_plain.log; tail -3 $O/ncu_scrfd_full.log
