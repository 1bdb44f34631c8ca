nproc; uptime
ITS=60 python tools/step_probe.py | awk '{ if ($10+0 > 15.5 || NR<4) print }' | head -30
echo "--- nice+nogc"
ITS=60 NICE=1 NOGC=1 python tools/step_probe.py | awk '{ if ($10+0 > 15.5 || NR<4) print }' | head -30
