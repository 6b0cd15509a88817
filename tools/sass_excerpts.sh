#!/bin/bash
# SASS evidence for the instructions DESIGN.md names (run after a build; needs cuobjdump, no GPU)
L=datmo_using_optical_flow_b200/lib/libdatmo_b200.so
fn() { cuobjdump -sass $L 2>/dev/null | awk -v pat="$1" '/Function : /{f = index($0, pat) > 0} f' ; }
echo "== k_pyr0_polyexp_t<uint8, 5, 4, blur>: packed f32x2 arithmetic and bulk-copy row stores (counts, then the first lines of each)"
fn k_pyr0_polyexp_tIhLi5ELi4ELb1 > /tmp/_p0.sass
for m in FFMA2 FADD2 FMUL2 UBLKCP; do echo "$m: $(grep -c "$m" /tmp/_p0.sass)"; done
grep -E "FFMA2|FADD2|FMUL2" /tmp/_p0.sass | head -12
grep -B3 -A2 "UBLKCP" /tmp/_p0.sass | head -24
echo
echo "== k_pyr_h_rows<7>: one output column of 32 rows — byte loads, 2^23 trick (VIADD 0x4b000000 + FADD -8388608), immediate taps, fp64 blend; no I2F"
fn k_pyr_h_rowsILi7 > /tmp/_ph.sass
echo "I2F: $(grep -c I2F /tmp/_ph.sass)   LDS.U8: $(grep -c 'LDS.U8' /tmp/_ph.sass)"
awk '/LDS.U8/{c++} c>=1 && n<46 {print; n++}' /tmp/_ph.sass
echo
echo "== k_flow_iter_xm<XmTile<46,320,2,2,2>>: the M phase's gathers (16-byte + 4-byte read-only loads through IMAD.WIDE addresses) and the L2 prefetch"
fn k_flow_iter_xmINS_6XmTileILi46 > /tmp/_xm.sass
for m in "LDG.E.128.CONSTANT" "LDG.E.CONSTANT" "LDG.E.64.CONSTANT" "CCTL" "LDS" "STS" "BAR.SYNC" "MUFU.RCP"; do echo "$m: $(grep -c "$m" /tmp/_xm.sass)"; done
grep -E "LDG.E.128.CONSTANT|CCTL|PREFETCH" /tmp/_xm.sass | head -14
