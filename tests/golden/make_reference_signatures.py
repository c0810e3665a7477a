"""Extracts, from the reference's own headers, the declarations of every member the sparsifier path touches, and commits them
as tests/golden/reference_signatures.json.  tests/test_reference_surface.py compares the shim headers
(ms_slam_b200/host/SlamShims.h, MapSparsification.h) with this fixture -- and, where /root/reference is present, the fixture
with the live headers -- so that the drop-in surface cannot drift unnoticed.  Run: python tests/golden/make_reference_signatures.py"""
import json, os, re, sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MSS_REFERENCE", "/root/reference")

# class -> members of the touch-set.  "*" after a name = data member
TOUCH = {
    "MapPoint": ["GetObservations", "Observations", "AddObservation", "UpdateObservation", "EraseObservation", "SetBadFlag", "isBad",
                 "GetMap", "GetIndexInKeyFrame", "mnId*", "nObs*", "mnMapSparsificationId*", "mnIndexForSparsification*"],
    "KeyFrame": ["GetNumberMPs", "AddMapPoint", "EraseMapPointMatch", "EraseBadDescriptor", "GetMapPointMatches", "GetMapPoint",
                 "GetMap", "UpdateCountInLocalMapping", "UpdateCountInTracking", "isNonLocal", "mnId*", "N*",
                 "mnMapSaprsificationId*", "mbSparsified*", "mnNonLocalKF*"],
    "Map": ["AddKeyFrame", "AddMapPoint", "EraseMapPoint", "AddSparsifiedMapPoint", "AddSparsifiedKeyFrame", "GetAllKeyFrames",
            "GetAllMapPoints", "MapPointsInMap", "SparsifiedMapPointsInMap", "GetAllSparsifiedKeyFrames", "SetIniertialBA2",
            "GetIniertialBA2"],
    "Atlas": ["GetCurrentMap", "GetAllKeyFrames"],
    "LoopClosing": ["InsertSparsifiedKeyFrame", "DeleteOutdatedInfo"],
    "MapSparsification": ["MapSparsification", "Run", "CheckNewKeyFrames", "InsertKeyFrame", "SetLoopClosing", "GetLastestKeyFrames",
                          "isStopped", "RequestStop", "Release", "RequestFinish", "isFinished", "mnMinNum*"],
}


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def class_body(text, cls):
    m = re.search(r"\bclass\s+" + cls + r"\b[^;{]*\{", text)
    if not m:
        return None
    i, depth = m.end(), 1
    while i < len(text) and depth:
        depth += {"{": 1, "}": -1}.get(text[i], 0)
        i += 1
    return text[m.end():i - 1]


def norm_type(t):
    t = re.sub(r"\bstd::", "", t)
    t = re.sub(r"\bORB_SLAM3::", "", t)
    t = re.sub(r"\s+", " ", t).strip()
    t = re.sub(r"\s*([<>,\*&])\s*", r"\1", t)
    t = re.sub(r"\b(static|virtual|inline|explicit)\b\s*", "", t)
    return t.strip()


def norm_param(p):
    p = p.strip()
    if not p:
        return None
    p = re.sub(r"=.*$", "", p).strip()                                   # default value
    m = re.match(r"^(.*?[\s\*&>])([A-Za-z_]\w*)$", p)                    # trailing identifier = the parameter's name
    if m and m.group(1).strip() and not re.fullmatch(r"(const|unsigned|long|int|short|signed)", m.group(2)):
        p = m.group(1)
    return norm_type(p)


def declarations(body, name, data_member):
    out = []
    if data_member:
        for m in re.finditer(r"([\w:<>,\s\*&]+?)\s+" + re.escape(name) + r"\s*(?:=[^;]*)?;", body):
            t = norm_type(m.group(1).split(";")[-1].split("}")[-1].split("{")[-1].split(":")[-1] if "::" not in m.group(1) else m.group(1).split(";")[-1])
            if t and t not in ("return", "else"):
                out.append(t)
        return sorted(set(out))
    for m in re.finditer(r"(?:^|[;{}:])\s*([\w:<>,\s\*&]*?)\s*\b" + re.escape(name) + r"\s*\(([^()]*)\)\s*(const)?", body):
        ret = norm_type(re.sub(r"\b(public|private|protected)\s*:", " ", m.group(1)))
        if ret in ("return", "else", "new", "delete"):
            continue
        params = [norm_param(p) for p in split_params(m.group(2))]
        out.append(f"{ret}({','.join(p for p in params if p)})")
    return sorted(set(out))


def split_params(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "<(":
            depth += 1
        elif ch in ">)":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur); cur = ""
        else:
            cur += ch
    parts.append(cur)
    return parts


def extract(header_dir, mapping=None):
    res = {}
    for cls, names in TOUCH.items():
        path = os.path.join(header_dir, (mapping or {}).get(cls, cls + ".h"))
        text = strip_comments(open(path).read())
        body = class_body(text, cls)
        if body is None:
            raise SystemExit(f"class {cls} not found in {path}")
        res[cls] = {}
        for n in names:
            dm = n.endswith("*")
            res[cls][n] = declarations(body, n.rstrip("*"), dm)
    return res


if __name__ == "__main__":
    ref = extract(os.path.join(REF, "include"))
    missing = [(c, n) for c, d in ref.items() for n, v in d.items() if not v]
    if missing:
        raise SystemExit(f"not found in the reference headers: {missing}")
    json.dump(ref, open(os.path.join(HERE, "reference_signatures.json"), "w"), indent=1, sort_keys=True)
    print("wrote", sum(len(d) for d in ref.values()), "members")
