"""CPU oracle for the qcdenoise noisy-circuit sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``qcdenoise_b200/`` may import this package: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs use it, and only as
the checker or as the timed CPU baseline -- never as the product path.

What it restates (file:line are relative to the reference checkout, nlaanait/qcdenoise):

* the reference's own Python for the path: gate lists (graph_circuit.py:69-173), noise-parameter draws
  (samplers.py:388-451, :554-727), counts -> probability vector (samplers.py:195-202, :37-68), stabilizer
  generators / sub-circuits (stabilizers.py:33-79, :104-168), diagonals / expectation values / witnesses
  (witnesses.py:40-144);
* the algorithm of the un-vendored third-party dependency the arithmetic lives in: qiskit-aer
  (requirements.txt:91-93 pins the stale metapackage ``qiskit==0.23.4``; the code needs Aer >= 0.8):
  gate-at-a-time complex128 statevector sweeps, one Kraus trajectory per shot, density-matrix evolution of
  the vectorised rho, ``sample_measure``; and qiskit-ignis 0.5.x ``mitigation.expectation_value``
  (requirements.txt:96) whose formula is recovered from, and pinned by, golden G1.

PARITY PIN STATUS: the counts -> diagonal -> <S> -> witness half is pinned by the reference's recorded
notebook outputs G1-G5 (tests/golden/witness_goldens.json, tests/test_oracle_goldens.py).  The
qiskit-Aer noisy-simulation boundary (noisy probability vectors) is **parity unpinned**: the reference's own
tests assert only shapes (tests/test_samplers.py:74,88), qiskit is neither vendored nor installable here, so
that half of the oracle is pinned only by analytic identities (CPTP, ideal graph-state closed form, gamma->0,
depolarizing contraction of Pauli expectations).
"""
