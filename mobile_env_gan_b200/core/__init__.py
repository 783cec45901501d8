"""Host-side mirror of the reference's ``mobile_env/core`` package: the same module and class names
(``base.MComCore``, ``channels``, ``movement``, ``arrival``, ``schedules``, ``utilities``, ``entities``,
``metrics``, ``logging``, ``util``) describing the batched CUDA simulator instead of executing the step
in Python.  ``views.EnvView`` gives a per-env snapshot with the reference's attribute names."""
