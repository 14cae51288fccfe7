#!/bin/bash
# profiles/r02_sass_summary.txt: which Blackwell instructions each object of the library holds
# (cuobjdump -sass on the objects smallpebble_b200.build compiled for sm_100a).
cd "$(dirname "$0")/../smallpebble_b200/_lib/obj" || exit 1
echo "# cuobjdump -sass of smallpebble_b200/_lib/obj/*.o (sm_100a), counts of the mnemonics that"
echo "# prove the execution model: UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = TMA"
echo "# tensor loads (IM2COL = im2col mode), LDTM = tcgen05.ld (TMEM -> registers), UTCBAR ="
echo "# tcgen05.commit, UCGABAR = cluster barrier, HMMA = warp-level mma.sync, LDGSTS = cp.async,"
echo "# ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents (programmatic dependent launch)"
for f in *.o; do
  echo "== $f"
  cuobjdump -sass "$f" 2>/dev/null | grep -oE "\b(UTCHMMA[A-Z0-9_.]*|UTMALDG[A-Z0-9_.]*|UTMASTG[A-Z0-9_.]*|LDTM[A-Z0-9_.]*|UTCBAR[A-Z0-9_.]*|UTCCP[A-Z0-9_.]*|HMMA[A-Z0-9_.]*|LDGSTS[A-Z0-9_.]*|ACQBULK|PREEXIT|UCGABAR_[A-Z]*)" | sort | uniq -c | sort -rn
done
