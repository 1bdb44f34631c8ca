f() { echo "== $*"; env "$@" ITS=40 python tools/step_probe.py | awk 'NR>3 { n++; s+=$15; if ($15+0>m) m=$15+0; t+=$11; if ($11+0>tm) tm=$11+0 } END { print "host knn mean", s/n, "max", m, "| dev total mean", t/n, "max", tm }'; }
f X=1
f PROFILE=1
f SAMPLER=1
f X=1
