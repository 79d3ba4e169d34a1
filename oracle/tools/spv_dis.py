#!/usr/bin/env python3
"""Minimal SPIR-V disassembler used ONCE, in the build container, to pin the arithmetic of the
reference's shipped physics shader (assets/shaders/wrach_physics_shaders.spv).

Test infrastructure only.  It reads /root/reference (which does not exist on the GPU box), so it is
never imported by tests, bench.py or the product; its *output* is committed as
tests/golden/spv_arith.json and checked by tests/test_oracle_golden.py.

Usage:  python oracle/tools/spv_dis.py /root/reference/assets/shaders/wrach_physics_shaders.spv \
            [--dump] [--json tests/golden/spv_arith.json]
"""
import json
import struct
import sys

OPS = {
    0: "Nop", 3: "Source", 4: "SourceExtension", 5: "Name", 6: "MemberName", 7: "String", 10: "Extension",
    11: "ExtInstImport", 12: "ExtInst", 14: "MemoryModel", 15: "EntryPoint", 16: "ExecutionMode",
    17: "Capability", 19: "TypeVoid", 20: "TypeBool", 21: "TypeInt", 22: "TypeFloat", 23: "TypeVector",
    28: "TypeArray", 29: "TypeRuntimeArray", 30: "TypeStruct", 32: "TypePointer", 33: "TypeFunction",
    41: "ConstantTrue", 42: "ConstantFalse", 43: "Constant", 44: "ConstantComposite", 46: "ConstantNull",
    54: "Function", 55: "FunctionParameter", 56: "FunctionEnd", 57: "FunctionCall", 59: "Variable",
    61: "Load", 62: "Store", 65: "AccessChain", 66: "InBoundsAccessChain", 68: "ArrayLength",
    71: "Decorate", 72: "MemberDecorate", 79: "VectorShuffle", 80: "CompositeConstruct",
    81: "CompositeExtract", 82: "CompositeInsert", 83: "CopyObject", 109: "ConvertFToU",
    110: "ConvertFToS", 111: "ConvertSToF", 112: "ConvertUToF", 113: "UConvert", 114: "SConvert",
    124: "Bitcast", 126: "SNegate", 127: "FNegate", 128: "IAdd", 129: "FAdd", 130: "ISub", 131: "FSub",
    132: "IMul", 133: "FMul", 134: "UDiv", 135: "SDiv", 136: "FDiv", 137: "UMod", 141: "FMod",
    142: "VectorTimesScalar", 148: "Dot", 164: "LogicalEqual", 166: "LogicalOr", 167: "LogicalAnd",
    168: "LogicalNot", 169: "Select", 170: "IEqual", 171: "INotEqual", 172: "UGreaterThan",
    173: "SGreaterThan", 174: "UGreaterThanEqual", 175: "SGreaterThanEqual", 176: "ULessThan",
    177: "SLessThan", 178: "ULessThanEqual", 179: "SLessThanEqual", 180: "FOrdEqual", 181: "FUnordEqual",
    182: "FOrdNotEqual", 183: "FUnordNotEqual", 184: "FOrdLessThan", 185: "FUnordLessThan",
    186: "FOrdGreaterThan", 187: "FUnordGreaterThan", 188: "FOrdLessThanEqual",
    189: "FUnordLessThanEqual", 190: "FOrdGreaterThanEqual", 191: "FUnordGreaterThanEqual",
    245: "Phi", 246: "LoopMerge", 247: "SelectionMerge", 248: "Label", 249: "Branch",
    250: "BranchConditional", 251: "Switch", 253: "Return", 254: "ReturnValue", 255: "Unreachable",
}
# GLSL.std.450 extended instruction numbers we care about
GLSL = {4: "FAbs", 8: "Floor", 31: "Sqrt", 32: "InverseSqrt", 37: "FMin", 40: "FMax", 43: "FClamp",
        46: "FMix", 50: "Fma", 66: "Length", 67: "Distance", 69: "Normalize"}
# opcodes that carry <result type, result id> as first two operands
HAS_TYPE_AND_RESULT = set(range(41, 47)) | {12, 54, 55, 57, 59, 61, 65, 66, 68, 79, 80, 81, 82, 83, 245} | \
    set(range(109, 115)) | {124} | set(range(126, 143)) | {148} | set(range(164, 192))
FLOAT_OPS = {"FAdd", "FSub", "FMul", "FDiv", "FNegate", "VectorTimesScalar", "Dot"}


def parse(path):
    raw = open(path, "rb").read()
    words = struct.unpack("<%dI" % (len(raw) // 4), raw)
    assert words[0] == 0x07230203, "not SPIR-V"
    header = {"version": "%d.%d" % ((words[1] >> 16) & 0xFF, (words[1] >> 8) & 0xFF),
              "bound": words[3], "words": len(words)}
    insts, i = [], 5
    while i < len(words):
        wc, op = words[i] >> 16, words[i] & 0xFFFF
        insts.append((op, list(words[i + 1:i + wc])))
        i += wc
    return header, insts


def fbits(w):
    return struct.unpack("<f", struct.pack("<I", w))[0]


def main(argv):
    path = argv[1]
    header, insts = parse(path)
    types, consts, defs = {}, {}, {}
    for op, a in insts:
        name = OPS.get(op, "Op%d" % op)
        if name == "TypeFloat":
            types[a[0]] = "f%d" % a[1]
        elif name == "TypeInt":
            types[a[0]] = ("i" if a[2] else "u") + str(a[1])
        elif name == "Constant":
            consts[a[1]] = fbits(a[2]) if types.get(a[0], "").startswith("f") else a[2]
        if op in HAS_TYPE_AND_RESULT and len(a) >= 2:
            defs[a[1]] = (name, a)

    def show(i):
        if i in consts:
            return "%%%d(=%r)" % (i, consts[i])
        return "%%%d" % i

    hist, glsl_hist, fmas, local_size, entry = {}, {}, [], None, None
    for op, a in insts:
        name = OPS.get(op, "Op%d" % op)
        if name == "EntryPoint":
            s = b"".join(struct.pack("<I", w) for w in a[2:])
            entry = s.split(b"\0")[0].decode()
        if name == "ExecutionMode" and a[1] == 17:
            local_size = a[2:5]
        if name in FLOAT_OPS:
            hist[name] = hist.get(name, 0) + 1
        if name == "ExtInst":
            g = GLSL.get(a[3], "glsl%d" % a[3])
            glsl_hist[g] = glsl_hist.get(g, 0) + 1
            if g == "Fma":
                fmas.append({"result": a[1], "a": a[4], "b": a[5], "c": a[6]})
        if "--dump" in argv:
            if op in HAS_TYPE_AND_RESULT and len(a) >= 2:
                extra = ""
                if name == "ExtInst":
                    extra = GLSL.get(a[3], "glsl%d" % a[3]) + " " + " ".join(show(x) for x in a[4:])
                elif name in ("CompositeExtract",):
                    extra = show(a[2]) + " idx " + " ".join(str(x) for x in a[3:])
                else:
                    extra = " ".join(show(x) for x in a[2:])
                print("%%%d = %s %s" % (a[1], name, extra))
            elif op >= 54:
                print("        %s %s" % (name, " ".join(show(x) for x in a)))

    def expr(i, depth=0):
        """Render the float expression tree feeding id i, stopping at loads / phis / non-float ops."""
        if i in consts:
            return repr(consts[i])
        if i not in defs or depth > 6:
            return "%%%d" % i
        name, a = defs[i]
        e = lambda k: expr(a[k], depth + 1)
        if name == "FAdd":
            return "(%s + %s)" % (e(2), e(3))
        if name == "FSub":
            return "(%s - %s)" % (e(2), e(3))
        if name == "FMul":
            return "(%s * %s)" % (e(2), e(3))
        if name == "FDiv":
            return "(%s / %s)" % (e(2), e(3))
        if name == "FNegate":
            return "-%s" % e(2)
        if name == "ExtInst":
            g = GLSL.get(a[3], "glsl%d" % a[3])
            return "%s(%s)" % (g.lower(), ", ".join(expr(x, depth + 1) for x in a[4:]))
        if name == "CompositeExtract":
            return "%s.%s" % (expr(a[2], depth + 1), "xyzw"[a[3]] if len(a) == 4 and a[3] < 4 else a[3:])
        if name == "Load":
            return "load%%%d" % a[2]
        return "%s%%%d" % (name, i)

    out = {
        "source": "assets/shaders/wrach_physics_shaders.spv",
        "header": header, "entry_point": entry, "local_size": local_size,
        "float_op_histogram": hist, "glsl_ext_histogram": glsl_hist,
        "fma_expressions": [expr(f["result"]) for f in fmas],
        "sqrt_expressions": [expr(i) for i, (n, a) in defs.items() if n == "ExtInst" and GLSL.get(a[3]) == "Sqrt"],
        "fdiv_expressions": [expr(i) for i, (n, a) in defs.items() if n == "FDiv"],
    }
    text = json.dumps(out, indent=1, sort_keys=True)
    if "--json" in argv:
        open(argv[argv.index("--json") + 1], "w").write(text + "\n")
    if "--dump" not in argv:
        print(text)


if __name__ == "__main__":
    main(sys.argv)
