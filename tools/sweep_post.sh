#!/bin/bash
cd "$(dirname "$0")/.." 2>/dev/null || cd /root/repo
for extra in "$@"; do
  make -C reina-vk_b200/csrc clean >/dev/null
  make -C reina-vk_b200/csrc -j8 EXTRA="$extra" >/dev/null 2>&1 || { echo "BUILD FAILED: $extra"; continue; }
  echo "EXTRA=[$extra] $(python tools/post_profile.py) | $(python tools/post_profile.py 3840 2160)"
done
make -C reina-vk_b200/csrc clean >/dev/null; make -C reina-vk_b200/csrc -j8 >/dev/null 2>&1
