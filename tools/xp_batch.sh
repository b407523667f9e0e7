#!/bin/bash
# experiment: which device-side stage limits the batch arm?  (library built with -DRQB_EXPERIMENTS)
sed -i 's/"-O3", "-march=x86-64-v3", "-std=c11"/"-O3", "-DRQB_EXPERIMENTS", "-march=x86-64-v3", "-std=c11"/' nanorq_b200/build.py
python -m nanorq_b200.build --force > /dev/null
python tools/pcie_probe3.py
for e in "XP_X=1" "XP_NO_SRC_D2H=1" "XP_NO_IMAGE_D2H=1" "XP_NO_COPYROWS=1" "XP_NO_RING_H2D=1" "XP_NO_PAYLOAD_H2D=1" "XP_NO_SRC_D2H=1 XP_NO_IMAGE_D2H=1" "XP_NO_RING_H2D=1 XP_NO_PAYLOAD_H2D=1" "XP_NO_SRC_D2H=1 XP_NO_IMAGE_D2H=1 XP_NO_RING_H2D=1 XP_NO_PAYLOAD_H2D=1"; do
  echo "== $e"; env XP_NOVERIFY=1 $e python tools/e2e_arms.py 4096 1280 0.1 0 118 20 2 2>&1 | grep batch
done
sed -i 's/"-O3", "-DRQB_EXPERIMENTS", /"-O3", /' nanorq_b200/build.py
python -m nanorq_b200.build --force > /dev/null
