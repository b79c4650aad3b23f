#!/bin/bash
# Where the real program spends its time with fft_engine = b200_eti (ODR_DABMOD_B200_TRACE), on a synthetic ETI file.
set -e
cd "$(dirname "$0")/.."
T=$(mktemp -d)
python - "$T" <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import dabmod_loader
dabmod_loader.load()
import importlib
eti = importlib.import_module("odr_dabmod_b200.eti")
fr = eti.synth_eti_range(1, eti.default_multiplex(), 0, 4000, seed=8)
for name, reps in (("in.eti", 10), ("in3.eti", 30)):
    with open(sys.argv[1] + "/" + name, "wb") as f:
        for _ in range(reps):
            f.write(fr.tobytes())
PY
for fmt in complexf u8; do
cat > $T/cfg.ini <<INI
[remotecontrol]
zmqctrl=0
telnet=0
[log]
syslog=0
[input]
transport=file
source=$T/in.eti
loop=0
[modulator]
fft_engine=b200_eti
gainmode=var
mode=1
rate=2048000
[firfilter]
enabled=0
[output]
output=file
[fileoutput]
format=$fmt
filename=/dev/null
INI
for d in 64 256; do
for f in in.eti in3.eti; do
  echo "== format $fmt depth $d file $f"
  sed -i "s#^source=.*#source=$T/$f#" $T/cfg.ini
  ( time ODR_DABMOD_B200_TRACE=1 ODR_DABMOD_B200_DEPTH=$d oracle/_ref/odr-dabmod-b200 $T/cfg.ini 2>&1 | grep -E "B200EtiChain|DAB frames" ) 2>&1 | grep -vE "^$|user|sys"
done
done
done
rm -rf $T
