O=gpurun_out/r02w; mkdir -p $O
timeout 300 python tools/umma16_timeline.py > $O/inkernel.txt 2>&1
python - <<'P'
import re
t=open('gpurun_out/r02w/inkernel.txt').read()
blocks=t.split('== launch')
for b in blocks[1:]:
    h=b.splitlines()[0]
    if 'grid.x 5 ' in h or 'grid.x 29 ' in h:
        print('== launch'+b[:3500])
P
