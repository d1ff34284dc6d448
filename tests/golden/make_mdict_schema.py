"""Freeze the variable names the reference drivers write to their .mat files.

Parses (ast, nothing is executed) the `mdict = {...}` literal of /root/reference/{SingleMassOscillator,VehicleSimulation,EMPS}_Simulation.py
and the file name passed to scipy.io.savemat, plus the names the drivers import from `src.*`, and writes tests/golden/mdict_schema.json.  Run in the build container (the
reference tree does not travel to the GPU box): python tests/golden/make_mdict_schema.py"""
import ast
import json
import os

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
out = {}
for name in ("SingleMassOscillator_Simulation", "VehicleSimulation_Simulation", "EMPS_Simulation"):
    tree = ast.parse(open(os.path.join(REF, name + ".py")).read())
    keys, target = None, None
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and isinstance(node.value, ast.Dict) and getattr(node.targets[0], "id", "") == "mdict":
            keys = [k.value for k in node.value.keys]
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "savemat":
            target = node.args[0].value
    imports = sorted({f"{node.module}.{a.name}" for node in ast.walk(tree) if isinstance(node, ast.ImportFrom) and node.module.startswith("src")
                      for a in node.names})
    out[name] = {"file": target, "keys": keys, "imports": imports}
json.dump(out, open(os.path.join(HERE, "mdict_schema.json"), "w"), indent=1)
print({k: (v["file"], len(v["keys"])) for k, v in out.items()})
