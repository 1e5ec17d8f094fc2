"""Per-kernel SASS mnemonic counts of libviettts_b200.so: the evidence that the tensor-core kernels are tcgen05 / TMEM /
bulk-copy code (UTC*MMA, LDTM/STTM, UBLKCP, UTCBAR) and that no legacy HMMA path exists.  Writes profiles/r2_sass_tcgen05.txt."""
import collections
import re
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
LIB = REPO / "viettts_b200" / "libviettts_b200.so"
PAT = {"UTCHMMA": r"\bUTCHMMA\b", "UTCHMMA.2CTA (cta_group::2)": r"\bUTCHMMA\.2CTA\b", "UTCBAR.2CTA.MULTICAST": r"\bUTCBAR\.2CTA\S*MULTICAST\b", "UTCQMMA/other UTC*MMA": r"\bUTC[A-Z]*MMA\b", "UTCBAR (tcgen05.commit)": r"\bUTCBAR\b", "LDTM (tcgen05.ld)": r"\bLDTM\b",
       "STTM (tcgen05.st)": r"\bSTTM\b", "UTCCP (tcgen05.cp)": r"\bUTCCP\b", "UBLKCP (cp.async.bulk)": r"\bUBLKCP\b", "UTMALDG/UTMASTG (tensor TMA)": r"\bUTMA(LDG|STG)\b",
       "SYNCS (mbarrier)": r"\bSYNCS\b", "HMMA (legacy mma.sync)": r"\bHMMA\b", "HGMMA (wgmma)": r"\bHGMMA\b", "FFMA": r"\bFFMA\b"}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for name, pat in PAT.items():
            if re.search(pat, line):
                funcs[cur][name] += 1
    dem = subprocess.run(["cu++filt"] + list(funcs), capture_output=True, text=True).stdout.splitlines()
    out = [f"SASS mnemonic counts per kernel of {LIB.name} (cuobjdump -sass, sm_100a).  tcgen05.mma -> UTCHMMA, tcgen05.commit -> UTCBAR,",
           "tcgen05.ld/st -> LDTM/STTM, cp.async.bulk -> UBLKCP, mbarrier -> SYNCS; HMMA / HGMMA would be the legacy tensor paths.", ""]
    tot = collections.Counter()
    for (mangled, cnt), name in zip(funcs.items(), dem if len(dem) == len(funcs) else list(funcs)):
        short = re.sub(r"\(anonymous namespace\)::", "", name)
        short = re.sub(r"\(int\)", "", short)
        short = re.sub(r"\(.*", "", short)
        tot.update(cnt)
        keys = [k for k in PAT if cnt[k] and k != "UTCQMMA/other UTC*MMA"]
        out.append(f"{short:70s} " + "  ".join(f"{k.split(' ')[0]}={cnt[k]}" for k in keys))
    out += ["", "TOTAL  " + "  ".join(f"{k.split(' ')[0]}={tot[k]}" for k in PAT if k != "UTCQMMA/other UTC*MMA"),
            f"legacy tensor instructions (HMMA / HGMMA): {tot['HMMA (legacy mma.sync)'] + tot['HGMMA (wgmma)']}"]
    src = "".join(p.read_text() for p in (REPO / "viettts_b200" / "csrc").glob("*.cu*"))
    out += ["", "PTX in the sources (inline asm):  " + "  ".join(f"{k}={len(re.findall(k, src))}" for k in
            (r"tcgen05\.mma", r"tcgen05\.ld", r"tcgen05\.st", r"tcgen05\.commit", r"tcgen05\.alloc", r"cp\.async\.bulk", r"mbarrier\."))]
    text = "\n".join(out) + "\n"
    (REPO / "profiles" / "r2_sass_tcgen05.txt").write_text(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
