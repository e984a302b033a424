"""Oracle-side inverter closures for a list of inverter specifications (oracle/cases.py)."""
from oracle import eval_prep_oracle as ep


def oracle_inverters(specs):
    out = []
    for sp in specs:
        if sp is None:
            out.append(None)
        elif sp[0] == 'resize':
            out.append(ep.resize_inverter(*sp[1:]))
        elif sp[0] == 'translate':
            out.append(ep.translate_inverter(*sp[1:]))
        else:
            out.append(lambda labels: labels)
    return out
