"""Synthetic MQT-bench-shaped OpenQASM 2.0 circuits.

The generators reproduce, for any width, the gate order of the 5-qubit MQT-bench files the
reference ships (``examples/ghz_5qubits.qasm:10-14``, ``examples/qft_5qubits.qasm:10-27``);
``tests/test_circuits.py`` checks n = 5 against those files line by line.  The text is fed to
the reference's own front end (qasm2ast -> ast2dqc_circuit -> partitioner -> compiler) by
``tests/golden/make_golden.py``; nothing here parses or compiles QASM.
"""

_HEADER = 'OPENQASM 2.0;\ninclude "qelib1.inc";\n'


def _tail(n, creg_lines):
    lines = ["barrier " + ",".join(f"q[{i}]" for i in range(n)) + ";"]
    lines += [f"measure q[{i}] -> meas[{i}];" for i in range(n)]
    return lines


def ghz_qasm(n):
    lines = [f"qreg q[{n}];", f"creg meas[{n}];", f"h q[{n - 1}];"]
    lines += [f"cx q[{i}],q[{i - 1}];" for i in range(n - 1, 0, -1)]
    return _HEADER + "\n".join(lines + _tail(n, 1)) + "\n"


def _pi_over(k):
    return "pi/" + str(2 ** k)


def qft_qasm(n):
    lines = [f"qreg q[{n}];", f"creg c[{n}];", f"creg meas[{n}];"]
    for i in range(n - 1, -1, -1):
        for j in range(n - 1, i, -1):
            lines.append(f"cp({_pi_over(j - i)}) q[{j}],q[{i}];")
        lines.append(f"h q[{i}];")
    for i in range(n // 2):
        lines.append(f"swap q[{i}],q[{n - 1 - i}];")
    return _HEADER + "\n".join(lines + _tail(n, 2)) + "\n"
