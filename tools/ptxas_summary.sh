#!/bin/sh
# (dev) registers / spills of every kernel in libramdsir.so:  sh tools/ptxas_summary.sh [file.so]
f="${1:-ramdsir_b200/libramdsir.so}"
cuobjdump --dump-resource-usage "$f" 2>/dev/null | awk '/Function/ {name=$2} /REG:/ {print name, $0}' | sed 's/Function //' | while read n rest; do echo "$(echo $n | sed 's/:$//' | c++filt | cut -c1-110) | $rest"; done
