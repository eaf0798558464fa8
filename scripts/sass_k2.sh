#!/bin/bash
# Development aid: static SASS of k2_wavefront from the built library: instruction count and opcode histogram.
cuobjdump -sass -fun _Z12k2_wavefront8K2Params minivideo_b200/libmvgpu.so | grep -E "^\s+/\*[0-9a-f]{4}\*/" > /tmp/k2.sass
wc -l < /tmp/k2.sass
sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?([A-Z0-9_]+).*/\2/' /tmp/k2.sass | sort | uniq -c | sort -rn | head -${1:-12}
