"""CPU oracle for the DDiffPG hot path (H1 sampler, H2 Q-ascent, H3 denoiser train step).

TEST INFRASTRUCTURE ONLY.  Nothing under ``ddiffpg_b200/`` may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs do, and there only as the checker or as the timed CPU baseline.

Parity pin: the reference ships no tests or golden vectors, and the DDPM
step arithmetic lives in third-party ``diffusers ^0.18.2`` (pyproject.toml:16) which is not
installed here.  The oracle is therefore pinned two ways:

* ``oracle/port.py`` (a functional restatement) is checked against outputs of the reference's
  OWN modules (``ddiffpg/models/diffusion_mlp.py``, ``ddiffpg/models/mlp.py``) imported by file
  path in the build container -- ``oracle/make_golden.py`` wrote ``tests/golden/*.npz``.
* ``oracle/ddpm.py`` (the restated diffusers scheduler) is checked against the reference's
  in-tree DDPM (``ddiffpg/models/baseline_models.py:59-185`` with
  ``baseline_helpers.cosine_beta_schedule``), also recorded in the golden files, and against
  the known-answer constants of SURVEY.md section 8(c).

The diffusers scheduler itself could not be executed ("parity unpinned" for that third-party
dependency alone); every other part of the oracle is pinned by reference-generated fixtures.
"""
