"""TEST INFRASTRUCTURE ONLY: rewrites the inline-PTX statements of gmrm_b200/csrc/kernels.cu into calls of the host
emulator (tests/emu/cuda_emu.h).  Every instruction form used by the step kernel is listed here; an unknown one is an
error, so a new asm statement in the product makes the emulation test fail loudly instead of silently skipping it."""
import re


def _balanced(text, i):
    """text[i] == '(' -> index just past the matching ')', skipping string literals."""
    depth, j = 0, i
    while j < len(text):
        c = text[j]
        if c == '"':
            j += 1
            while text[j] != '"':
                j += 2 if text[j] == "\\" else 1
        elif c == "(":
            depth += 1
        elif c == ")":
            depth -= 1
            if depth == 0:
                return j + 1
        j += 1
    raise ValueError("unbalanced asm statement")


def _split_top(s, sep):
    out, depth, cur, j = [], 0, "", 0
    while j < len(s):
        c = s[j]
        if c == '"':
            k = j + 1
            while s[k] != '"':
                k += 2 if s[k] == "\\" else 1
            cur += s[j:k + 1]
            j = k + 1
            continue
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if depth == 0 and s.startswith(sep, j) and not (sep == ":" and (s.startswith("::", j) and False)):
            out.append(cur)
            cur = ""
            j += len(sep)
            continue
        cur += c
        j += 1
    out.append(cur)
    return out


def _operands(section):
    """'"=r"(v), "l"(p)' -> ['v', 'p']"""
    ops = []
    for item in _split_top(section, ","):
        item = item.strip()
        if not item:
            continue
        m = re.match(r'"[^"]*"\s*\(', item)
        assert m, item
        ops.append(item[m.end():-1].strip())
    return ops


def _translate(instr, outs, ins):
    i = instr.strip()
    if i == "":                                              # empty statement: an optimisation barrier for the compiler only
        return "(void)0;"
    if i.startswith("createpolicy."):                        # L2 eviction policy of the genotype stream: no meaning on the host
        return f"{outs[0]} = 0;"
    if i.startswith("ld.global.nc.L1::no_allocate.L2::cache_hint.u32"):
        return f"{outs[0]} = *reinterpret_cast<const uint32_t*>({ins[0]});"
    if i.startswith("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32"):
        return ("{ const uint4 emu_t_ = *reinterpret_cast<const uint4*>(" + ins[0] + "); "
                f"{outs[0]} = emu_t_.x; {outs[1]} = emu_t_.y; {outs[2]} = emu_t_.z; {outs[3]} = emu_t_.w; }}")
    if i.startswith("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32"):
        return ("{ const uint32_t* emu_t_ = reinterpret_cast<const uint32_t*>(" + ins[0] + "); "
                f"{outs[0]} = emu_t_[0]; {outs[1]} = emu_t_[1]; }}")
    if i.startswith("ld.global.nc.L1::no_allocate.u32"):
        return f"{outs[0]} = *reinterpret_cast<const uint32_t*>({ins[0]});"
    if i.startswith("ld.global.nc.L1::no_allocate.v4.u32"):
        return ("{ const uint4 emu_t_ = *reinterpret_cast<const uint4*>(" + ins[0] + "); "
                f"{outs[0]} = emu_t_.x; {outs[1]} = emu_t_.y; {outs[2]} = emu_t_.z; {outs[3]} = emu_t_.w; }}")
    if i.startswith("ld.global.nc.L1::no_allocate.v2.u32"):
        return ("{ const uint32_t* emu_t_ = reinterpret_cast<const uint32_t*>(" + ins[0] + "); "
                f"{outs[0]} = emu_t_[0]; {outs[1]} = emu_t_[1]; }}")
    if i.startswith("{") and "prmt.b32 lo, %1, 0, %2" in i:    # direct rows: low word of the double <- one byte of m, high word kept
        return ("{ uint64_t emu_r_; memcpy(&emu_r_, &(" + outs[0] + "), 8); "
                f"emu_r_ = (emu_r_ & 0xffffffff00000000ull) | emu_prmt({ins[0]}, 0u, {ins[1]}); memcpy(&({outs[0]}), &emu_r_, 8); }}")
    if i.startswith("st.global.v4.f64"):                     # 32-byte store, word by word (no atomicity of the whole is assumed)
        return ("{ volatile double* emu_t_ = reinterpret_cast<volatile double*>(" + ins[0] + "); "
                f"emu_t_[0] = {ins[1]}; emu_t_[1] = {ins[2]}; emu_t_[2] = {ins[3]}; emu_t_[3] = {ins[4]}; }}")
    if i.startswith("ld.volatile.global.v4.u64"):
        return ("{ const volatile unsigned long long* emu_t_ = reinterpret_cast<const volatile unsigned long long*>(" + ins[0] + "); "
                f"{outs[0]} = emu_t_[0]; {outs[1]} = emu_t_[1]; {outs[2]} = emu_t_[2]; {outs[3]} = emu_t_[3]; }}")
    if i.startswith("ld.shared.f64 %0, [%1+%2]"):
        return f"{outs[0]} = emu_lds<double>(({ins[0]}) + (uint32_t)({ins[1]}));"
    if i.startswith("ld.shared.u32 %0, [%1+%2]"):
        return f"{outs[0]} = emu_lds<uint32_t>(({ins[0]}) + (uint32_t)({ins[1]}));"
    if i.startswith("ld.shared.u32 %0, [%1]"):
        return f"{outs[0]} = emu_lds<uint32_t>({ins[0]});"
    if i.startswith("st.shared.f64 [%0], %1"):
        return f"emu_sts<double>({ins[0]}, {ins[1]});"
    if i.startswith("prmt.b32"):
        return f"{outs[0]} = emu_prmt({ins[0]}, {ins[1]}, {ins[2]});"
    if i.startswith("prefetch.global.L2") or i.startswith("prefetch.global.L1") or i.startswith("cp.async.bulk.prefetch.L2"):
        return "(void)0;"
    if i.startswith("griddepcontrol."):                      # programmatic dependent launch: launches are serial on the host
        return "(void)0;"
    raise ValueError(f"asm_to_host: no host form for PTX instruction: {instr!r}")


def rewrite(text):
    out, pos, n = "", 0, 0
    for m in re.finditer(r"\basm\s*(volatile)?\s*\(", text):
        if m.start() < pos:
            continue
        end = _balanced(text, m.end() - 1)
        body = text[m.end():end - 1]
        semi = end
        while text[semi] in " \t":
            semi += 1
        assert text[semi] == ";", text[m.start():end + 5]
        parts = _split_top(body, ":")
        instr = parts[0].strip()
        assert instr.startswith('"') and instr.endswith('"'), instr
        # "a" : outs : ins : clobbers   (an empty output list is written "::")
        outs = _operands(parts[1]) if len(parts) > 1 else []
        ins = _operands(parts[2]) if len(parts) > 2 else []
        out += text[pos:m.start()] + _translate(instr[1:-1], outs, ins)
        pos = semi + 1
        n += 1
    return out + text[pos:], n
