#!/bin/bash
# K3 build variants side by side (libaewave_<tag>.so next to libaewave.so): screen + median time per 35,800 windows
cd /root/repo/ae_waveforms_b200/csrc
cp libaewave.so /tmp/libaewave_base.so
for tag in base "$@"; do
  [ "$tag" = base ] && cp /tmp/libaewave_base.so libaewave.so || cp libaewave_$tag.so libaewave.so
  echo "== $tag"; (cd /root/repo; NO_TICKS=1 python profiles/k3_phase_timing.py 2>&1 | tail -3 | cut -c1-120)
done
cp /tmp/libaewave_base.so libaewave.so
