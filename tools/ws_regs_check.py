"""Register budget of gram_ws_kernel after its role split (setmaxnreg): the helper warps give registers to the DMMA warps,
so no instruction a helper can reach may name a register at or above REG_HLP, and none at all one at or above REG_MMA.
ptxas reports only the launch-time count (128), so this reads the SASS: the code between USETMAXREG.TRY_ALLOC (DMMA branch
entered) and USETMAXREG.DEALLOC (helper branch entered) is the DMMA branch; the prologue and everything after it up to the
first out-of-line subroutine is helper-reachable; a subroutine (the slow paths of the FP64 division / square root, CALL
target .. RET) counts for every branch that calls it.
usage: python tools/ws_regs_check.py [path/to/assemble_fused.o]   (exit code 1 on a violation)"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def check(obj):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    out, cur, rows = [], None, []

    def flush():
        if not (cur and "gram_ws_kernel" in cur and rows):
            return
        marks = [(a, i) for a, i in rows if "USETMAXREG" in i]
        assert len(marks) == 2 and "TRY_ALLOC" in marks[0][1] and "DEALLOC" in marks[1][1], (cur, marks)
        reg_mma = int(marks[0][1].split(",")[-1], 16)
        reg_hlp = int(marks[1][1].split()[-1], 16)
        # out-of-line subroutines: [CALL target, next RET]; the main body ends at the first of them
        entries = sorted({int(i.split()[-1], 16) for a, i in rows if i.startswith("CALL") or " CALL" in i})
        rets = sorted(a for a, i in rows if "RET." in i)
        subs = [(e, min(r for r in rets if r >= e) + 16) for e in entries]
        body_end = min([e for e, _ in subs] + [rows[-1][0] + 16])
        regs = lambda lo, hi: [int(r) for a, i in rows if lo <= a < hi for r in re.findall(r"\bR(\d+)\b", i)] + [-1]
        in_mma = lambda a: marks[0][0] <= a < marks[1][0]
        m_mma = max(regs(marks[0][0], marks[1][0]))
        m_hlp = max(max(regs(0, marks[0][0])), max(regs(marks[1][0], body_end)))
        for e, end in subs:  # a subroutine counts for every branch that calls it
            callers = [a for a, i in rows if ("CALL" in i) and int(i.split()[-1], 16) == e]
            m = max(regs(e, end))
            if any(in_mma(a) for a in callers):
                m_mma = max(m_mma, m)
            if any(not in_mma(a) for a in callers):
                m_hlp = max(m_hlp, m)
        out.append((re.search(r"gram_ws_kernelILi(\d)ELi(\d)ELi(\d)", cur).groups(), reg_mma, m_mma, reg_hlp, m_hlp))

    for line in sass.splitlines():
        if "Function :" in line:
            flush()
            cur, rows = line, []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m:
            rows.append((int(m.group(1), 16), m.group(2)))
    flush()
    return out


if __name__ == "__main__":
    obj = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dune-iga_b200", "csrc", "_build", "assemble_fused.o")
    bad = False
    for inst, reg_mma, m_mma, reg_hlp, m_hlp in check(obj):
        ok = m_mma < reg_mma and m_hlp < reg_hlp and reg_mma + reg_hlp <= 256
        bad |= not ok
        print("gram_ws_kernel<%s,%s,%s>: DMMA branch R%d of %d, helper-reachable code R%d of %d  %s" %
              (*inst, m_mma, reg_mma, m_hlp, reg_hlp, "ok" if ok else "VIOLATION"))
    sys.exit(1 if bad else 0)
