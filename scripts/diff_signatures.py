"""Signature-level parity with the reference: every public function, class constructor and public method that exists in both
packages must accept the reference's parameter names (keyword callers) in the reference's positional order (positional
callers), and must not add required parameters.  ``python scripts/diff_signatures.py`` prints the differences;
``tests/test_reference_surface.py`` pins the list to the deliberate ones (the query router and the replicator take a ``send``
callable where the reference takes a libp2p host / stream: different transport).

Needs the reference importable: ``PYTHONPATH=baseline/shims:baseline/_ref:.``"""
import importlib
import inspect
import os


def collect():
    mods=[]
    for root,_,fs in os.walk('/root/reference/infomesh'):
        for f in fs:
            if f.endswith('.py'):
                rel=os.path.relpath(os.path.join(root,f),'/root/reference/infomesh')[:-3].replace('/','.')
                mods.append(rel[:-9] if rel.endswith('.__init__') else rel)
    def params(fn):
        try: sig=inspect.signature(fn)
        except (TypeError,ValueError): return None
        return [(p.name,p.kind.name, p.default is not inspect._empty) for p in sig.parameters.values() if p.name not in ("self","cls")]
    issues=[]
    for m in sorted(set(mods)):
        if m in ("__init__",""): continue
        try: a=importlib.import_module("infomesh."+m); b=importlib.import_module("infomesh_b200."+m)
        except Exception as e: continue
        for name,obj in vars(a).items():
            if name.startswith("_") or getattr(obj,"__module__",None)!=a.__name__: continue
            other=getattr(b,name,None)
            if other is None: continue
            pairs=[]
            if inspect.isfunction(obj): pairs.append((name,obj,other))
            elif inspect.isclass(obj):
                pairs.append((name+".__init__",obj.__init__,getattr(other,"__init__",None)))
                for mn,mo in vars(obj).items():
                    if mn.startswith("_") or not inspect.isfunction(mo): continue
                    om=getattr(other,mn,None)
                    if om is not None and callable(om): pairs.append((f"{name}.{mn}",mo,om))
            for label,f1,f2 in pairs:
                if f2 is None: continue
                p1,p2=params(f1),params(f2)
                if p1 is None or p2 is None: continue
                n2=[x[0] for x in p2]; var_kw=any(k=="VAR_KEYWORD" for _,k,_ in p2); var_pos=any(k=="VAR_POSITIONAL" for _,k,_ in p2)
                missing=[x[0] for x in p1 if x[0] not in n2 and x[1] not in ("VAR_KEYWORD","VAR_POSITIONAL") and not var_kw]
                # positional order: reference positional params must appear in the same relative order at the same index
                pos1=[x[0] for x in p1 if x[1] in ("POSITIONAL_ONLY","POSITIONAL_OR_KEYWORD")]
                pos2=[x[0] for x in p2 if x[1] in ("POSITIONAL_ONLY","POSITIONAL_OR_KEYWORD")]
                order_bad = (not var_pos) and any(i>=len(pos2) or pos2[i]!=n for i,n in enumerate(pos1) if n in n2)
                # new required params in ours
                req_new=[x[0] for x in p2 if not x[2] and x[1] in ("POSITIONAL_OR_KEYWORD","KEYWORD_ONLY") and x[0] not in [y[0] for y in p1]]
                if missing or order_bad or req_new:
                    issues.append((m,label,missing,order_bad,req_new,pos1,pos2))
    return issues


if __name__ == "__main__":
    found = collect()
    for mod, label, missing, order_bad, new_required, ref_pos, our_pos in found:
        print(f"{mod}:{label}  missing={missing} order_differs={order_bad} new_required={new_required}\n     ref{ref_pos}\n     our{our_pos}")
    print(len(found), "signature differences")
