"""Empty stand-in so that the reference's ``agent/r2d2.py`` (which does ``import gym`` and never uses it on the
learner path) can be imported unmodified -- TEST INFRASTRUCTURE ONLY (see ../tensorflow/__init__.py)."""
