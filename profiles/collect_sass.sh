#!/bin/bash
# Regenerates profiles/sass_summary.txt from the in-tree sm_100a libraries (no GPU needed).
cd "$(dirname "$0")/.."
out=profiles/sass_summary.txt
: > $out
for so in gemm coll ops; do
  f=batch_shipyard_b200/_native/libshipyard_$so.so
  echo "== $f ($(cuobjdump -lelf $f 2>/dev/null | grep -c sm_100a) sm_100a cubins)" >> $out
  cuobjdump -sass $f 2>/dev/null | grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" | awk '{print $2}' | sort | uniq -c | sort -rn \
    | grep -E "UTCHMMA|UTCBAR|LDTM|STTM|UTMALDG|UTMASTG|UTCATOMSWS|SYNCS|REDG|LDGMC|\.MC|MULTIMEM|ATOMG|UBLKCP|\.STRONG\.SYS|HMMA|DFMA" >> $out
  echo >> $out
done
echo "-- per-kernel registers / smem (cuobjdump -res-usage)" >> $out
for so in gemm coll ops; do
  cuobjdump -res-usage batch_shipyard_b200/_native/libshipyard_$so.so 2>/dev/null | grep -A1 "Function" | grep -vE "^--" | paste - - | sed -E 's/ +/ /g' | cut -c1-260 >> $out
done

# per-instantiation mnemonic counts of the halo-load kernels (tcgen05 MMA, tiled 4-D TMA halo box, TMEM loads)
python3 - >> $out <<'PY'
import re, subprocess
txt = subprocess.run(["cuobjdump", "-sass", "batch_shipyard_b200/_native/libshipyard_gemm.so"], stdout=subprocess.PIPE, text=True).stdout
print("\n-- halo-load kernels: tcgen05 / TMA mnemonics per instantiation (validated on B200: conv3x3_halo_kernel<BN, stats, dgrad, pair, 0, 0>)")
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    if "halo" not in name:
        continue
    cnt = {m: len(re.findall(r"\b" + re.escape(m), f)) for m in ("UTCHMMA", "UTMALDG.4D", "UTMALDG.2D", "UTMALDG.3D", "UTCBAR", "LDTM", "REDG", "SYNCS.PHASECHK")}
    short = re.sub(r"Ev14CUtensorMap.*", "", re.sub(r"^_Z\d+", "", name))
    print(f"{short:60s} " + " ".join(f"{k}={v}" for k, v in cnt.items() if v))
PY

# round-2 kernels: stem (32-byte-swizzle tcgen05 operands), scatter-epilogue dgrad, multi-block LL / broadcast scatter+all-gather collectives
python3 - >> $out <<'PY'
import re, subprocess
def funcs(lib):
    txt = subprocess.run(["cuobjdump", "-sass", f"batch_shipyard_b200/_native/{lib}"], stdout=subprocess.PIPE, text=True).stdout
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        yield f.split("\n", 1)[0].strip(), f
print("\n-- round-2 kernels: mnemonics per function")
for name, f in funcs("libshipyard_gemm.so"):
    if "stem_s2d" in name:
        cnt = {m: len(re.findall(r"\b" + re.escape(m), f)) for m in ("UTCHMMA", "UTMALDG.4D", "UTMASTG.4D", "UTCBAR", "LDTM", "REDG", "SYNCS.PHASECHK", "UTCATOMSWS")}
        print(f"{re.sub(r'^_Z[0-9]+', '', name)[:48]:50s} " + " ".join(f"{k}={v}" for k, v in cnt.items() if v))
for name, f in funcs("libshipyard_coll.so"):
    if any(k in name for k in ("k_lm_k", "k_broadcast_sag_k", "k_mailbox_k", "k_twoshot_nvls", "k_fused_sgd_k")):
        cnt = {m: len(re.findall(r"\b" + re.escape(m), f)) for m in ("LDGMC", "STG.E.128.STRONG.SYS", "STG.E.STRONG.SYS", "LDG.E.128.STRONG.SYS", "LDG.E.STRONG.SYS", "REDG", "ST.E.128", "MEMBAR.SC.SYS", "MEMBAR.ALL.SYS", "ATOMG")}
        mc = len(re.findall(r"\bST[G]?\.E[A-Z0-9.]*MC|MULTIMEM|\.MC\b", f))
        print(f"{re.sub(r'^_Z[0-9]+', '', name)[:48]:50s} " + " ".join(f"{k}={v}" for k, v in cnt.items() if v) + (f" multicast-stores={mc}" if mc else ""))
PY
