"""CPU oracle for the BPR-MF hot path of yoongi0428/RecSys_PyTorch.

TEST INFRASTRUCTURE ONLY.  Nothing under ``recsys_pytorch_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may.  The product path is CUDA
only and raises when its extension is missing.

Contents
--------
``bpr_oracle.py``   numpy restatement (fp32) of the reference's training-step and
                    scoring arithmetic (models/MF.py, models/LightGCN.py).
``eval_oracle.c``   plain-C restatement of the reference's native evaluation layer
                    (evaluation/backend/cython/include/{func,holdout,loo}.h).
``ref_eval_shim.cpp``  extern "C" doorway onto the *unmodified* reference headers,
                    compiled from where they lie in /root/reference into
                    ``oracle/_ref/`` (never copied into this repo).
``ref_harness.py``  imports the Python reference from /root/reference with the
                    three arithmetic-neutral shims of SURVEY.md section 8(c).
``make_golden.py``  runs the reference here and writes ``tests/golden/*.npz``.

Parity pin: upstream ships no tests or golden vectors ("parity unpinned"
upstream); the pin adopted here is *outputs of the reference itself executed in
the build container* (torch 2.11.0 CPU fp32, 1 thread), committed under
``tests/golden/`` together with ``make_golden.py`` that produced them.
"""
