"""NumPy/SciPy interpreter of the graph bytecode (pbl_graph_instr, include/probabilit_b200.h).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  It lets the CPU test-suite check the host
compiler of probabilit_b200/modeling.py (column assignment, evaluation order, dtype rules, slot
allocation) against the reference's golden vectors without a GPU, and gives the GPU tests an
instruction-level checker for csrc/graph.cu.  Every opcode is evaluated with the NumPy / SciPy
call the reference itself makes for that node type (modeling.py:795-812, :962-1169).
"""
import numpy as np
from scipy import stats

OPS = dict(
    NOP=0, LOAD=1, STORE=2, CHECK=3, MOV=4, UNIFORM=5,
    PPF_NORM=16, PPF_UNIFORM=17, PPF_EXPON=18, PPF_TRIANG=19, PPF_GAMMA=20, PPF_LOGNORM=21,
    PPF_POISSON=22, PPF_BINOM=23, PPF_BERNOULLI=24, TABLE_INTERP=25, TABLE_SEARCH=26, TABLE_QUANTILE=27,
    PPF_BETA=28, PPF_TRUNCNORM=29,
    ADD=32, MUL=33, SUB=34, DIV=35, POW=36, FLOORDIV=37, MOD=38, MAX=39, MIN=40, ATAN2=41, LT=42, LE=43,
    GT=44, GE=45, EQ=46, NE=47, AND=48, OR=49, ISCLOSE=50,
    NEG=64, ABS=65, LOG=66, EXP=67, FLOOR=68, CEIL=69, SIGN=70, SQRT=71, SQUARE=72, LOG10=73, SIN=74,
    COS=75, TAN=76, ASIN=77, ACOS=78, ATAN=79, SINH=80, COSH=81, TANH=82, ASINH=83, ACOSH=84, ATANH=85,
    NOT=86, LOOKUP=87,
)
QUANTILE_METHODS = ["linear", "lower", "higher", "nearest", "midpoint", "closest_observation"]
NAME = {v: k for k, v in OPS.items()}

BINARY = {
    "ADD": np.add, "MUL": np.multiply, "SUB": np.subtract, "DIV": np.true_divide, "POW": np.power,
    "FLOORDIV": np.floor_divide, "MOD": np.mod, "MAX": np.maximum, "MIN": np.minimum, "ATAN2": np.arctan2,
    "LT": np.less, "LE": np.less_equal, "GT": np.greater, "GE": np.greater_equal, "EQ": np.equal,
    "NE": np.not_equal, "AND": np.logical_and, "OR": np.logical_or, "ISCLOSE": np.isclose,
}
UNARY = {
    "NEG": np.negative, "ABS": np.abs, "LOG": np.log, "EXP": np.exp, "FLOOR": np.floor, "CEIL": np.ceil,
    "SIGN": np.sign, "SQRT": np.sqrt, "SQUARE": np.square, "LOG10": np.log10, "SIN": np.sin, "COS": np.cos,
    "TAN": np.tan, "ASIN": np.arcsin, "ACOS": np.arccos, "ATAN": np.arctan, "SINH": np.sinh, "COSH": np.cosh,
    "TANH": np.tanh, "ASINH": np.arcsinh, "ACOSH": np.arccosh, "ATANH": np.arctanh,
    "NOT": lambda a: a == 0,
}


def ppf(name, q, p):
    """scipy.stats.<distr>(...).ppf(q) with the operand order of the PBL_PPF_* opcodes."""
    with np.errstate(all="ignore"):
        if name == "PPF_NORM":
            return stats.norm(loc=p[0], scale=p[1]).ppf(q)
        if name == "PPF_UNIFORM":
            return stats.uniform(loc=p[0], scale=p[1]).ppf(q)
        if name == "PPF_EXPON":
            return stats.expon(loc=p[0], scale=p[1]).ppf(q)
        if name == "PPF_TRIANG":
            return stats.triang(p[0], loc=p[1], scale=p[2]).ppf(q)
        if name == "PPF_GAMMA":
            return stats.gamma(p[0], loc=p[1], scale=p[2]).ppf(q)
        if name == "PPF_LOGNORM":
            return stats.lognorm(p[0], loc=p[1], scale=p[2]).ppf(q)
        if name == "PPF_POISSON":
            return stats.poisson(p[0], loc=p[1]).ppf(q)
        if name == "PPF_BINOM":
            return stats.binom(p[0], p[1], loc=p[2]).ppf(q)
        if name == "PPF_BERNOULLI":
            return stats.bernoulli(p[0], loc=p[1]).ppf(q)
        if name == "PPF_BETA":  # operands (a, b, scale); loc is added by the next instruction
            return stats.beta(p[0], p[1], scale=p[2]).ppf(q)
        if name == "PPF_TRUNCNORM":
            return stats.truncnorm(p[0], p[1], scale=p[2]).ppf(q)
    raise ValueError(name)


ACC0, ACC1, NO_WRITE = 0x1000, 0x2000, 0x4000  # internal flags of the device form (csrc/graph.cu)


def run(program, n_slots, n, inputs, outputs, uniform=None, dev_ops=None):
    """program: sequence of objects with .op .dst .src[4] .imm[4]; inputs / outputs: lists of
    length-n float64 arrays (outputs are written in place).  Values are float64 like on the device
    (booleans as 0.0 / 1.0).  Returns the smallest failing CHECK tag or -1.

    dev_ops (optional): the op words of the library's DEVICE FORM of the program
    (pbl_graph_debug_translate): operands flagged ACC0 / ACC1 are taken from the previous instruction's
    result instead of a slot, and a NO_WRITE instruction writes no slot; slots start out as None, so a read
    of a value the translation pass wrongly dropped surfaces as an error here."""
    slots = [None] * max(n_slots, 1)
    bad = -1
    acc = None
    dflags = 0

    def operand(ins, i):
        if (i == 0 and dflags & ACC0) or (i == 1 and dflags & ACC1):
            return acc
        s = ins.src[i]
        return slots[s] if s >= 0 else np.full(n, ins.imm[i])

    def write(dst, value):
        nonlocal acc
        acc = value
        if not dflags & NO_WRITE:  # (the device form drops never-materialised slots: dst is meaningless then)
            slots[dst] = value

    with np.errstate(all="ignore"):
        for pc, ins in enumerate(program):
            dflags = dev_ops[pc] if dev_ops is not None else 0
            name = NAME[ins.op & 0xFF]
            flags, dst, tag, oidx = ins.op, ins.dst & 0xFF, (ins.dst >> 8) & 0xFFF, (ins.dst & 0xFFFFFFFF) >> 20

            def finish(value):
                nonlocal bad
                write(dst, value)
                if flags & 0x400 and not np.all(np.isfinite(value)):
                    bad = tag if bad < 0 else min(bad, tag)
                if flags & 0x800:
                    outputs[oidx][:] = value

            if name == "LOAD":
                write(ins.dst & 0xFF, np.array(inputs[ins.src[0]], dtype=np.float64))
            elif name == "STORE":
                outputs[ins.src[1]][:] = slots[ins.src[0]]
            elif name == "CHECK":
                if not np.all(np.isfinite(slots[ins.src[0]])):
                    bad = ins.src[1] if bad < 0 else min(bad, ins.src[1])
            elif name == "MOV":
                finish(operand(ins, 0).copy())
            elif name == "UNIFORM":
                write(ins.dst & 0xFF, uniform(ins.src[0]))
            elif name.startswith("TABLE_"):
                if flags & 0x100:
                    q = np.array(inputs[ins.src[0]], dtype=np.float64)
                elif flags & 0x200:
                    q = uniform(ins.src[0])
                else:
                    q = operand(ins, 0)
                m = int(ins.imm[1])
                tab = np.asarray(inputs[ins.src[1]])
                if name == "TABLE_INTERP":  # CumulativeDistribution._sample, modeling.py:880-882
                    out = np.interp(x=q, xp=tab[:m], fp=tab[m:2 * m])
                elif name == "TABLE_SEARCH":  # DiscreteDistribution._sample, modeling.py:910-912
                    out = np.searchsorted(tab[:m], v=q, side="right").astype(np.float64)
                else:  # EmpiricalDistribution._sample, modeling.py:841-842
                    out = np.quantile(a=tab[:m], q=q, method=QUANTILE_METHODS[int(ins.imm[2])])
                finish(np.asarray(out, dtype=np.float64))
            elif name == "LOOKUP":
                tab = np.asarray(inputs[ins.src[1]])[: int(ins.imm[1])]
                finish(tab[operand(ins, 0).astype(np.intp)].astype(np.float64))
            elif name.startswith("PPF_"):
                if flags & 0x100:
                    q = np.array(inputs[ins.src[0]], dtype=np.float64)
                elif flags & 0x200:
                    q = uniform(ins.src[0])
                else:
                    q = operand(ins, 0)
                finish(np.asarray(ppf(name, q, [operand(ins, i) for i in (1, 2, 3)]), dtype=np.float64))
            elif name in BINARY:
                finish(np.asarray(BINARY[name](operand(ins, 0), operand(ins, 1)), dtype=np.float64))
            elif name in UNARY:
                finish(np.asarray(UNARY[name](operand(ins, 0)), dtype=np.float64))
            elif name != "NOP":
                raise ValueError(f"unknown opcode {ins.op}")
    return bad
