"""CPU: the procedural scene generators of the BASELINE configs are deterministic and have the stated shape."""
import numpy as np

from mray_b200 import scenes


def test_config4_instanced_field():
    f, g = scenes.instanced_field(), scenes.instanced_field()
    assert len(f["meshes"]) == 12 and len(f["instances"]) == 1000 + 1 + 16
    sizes = [m[1].shape[0] for m in f["meshes"][:10]]
    assert 900 <= min(sizes) <= 1100 and 95_000 <= max(sizes) <= 105_000          # 1 K .. 100 K triangles
    assert abs(f["triangles_instanced"] - 10_000_000) < 500_000
    mats = [m for _, _, m in f["instances"][:1000]]
    assert mats == [k % 64 for k in range(1000)]                                   # round-robin per instance
    assert sum(1 for _, _, m in f["instances"] if m < 0) == 16                    # light panels
    for (ma, Ta, ka), (mb, Tb, kb) in zip(f["instances"], g["instances"]):
        assert ma == mb and ka == kb and (Ta is None) == (Tb is None) and (Ta is None or np.array_equal(Ta, Tb))
    for (pa, ia), (pb, ib) in zip(f["meshes"], g["meshes"]):
        assert np.array_equal(pa, pb) and np.array_equal(ia, ib)
        assert ia.max() < pa.shape[0]
    # every instance transform is a similarity (rotation * uniform scale + translation)
    for _, T, _ in f["instances"][:50]:
        R = T[:, :3]; s2 = (R @ R.T)[0, 0]
        assert np.allclose(R @ R.T, s2 * np.eye(3), atol=1e-9)


def test_cornell_variants():
    c = scenes.cornell_box()
    assert c["indices"].shape == (36, 3) and c["positions"].shape == (72, 3)
    uvs, textures, at = scenes.cornell_textures()
    assert uvs.shape == (72, 2) and uvs.min() < 0 and uvs.max() > 1              # wrap-around and negative texels occur
    assert textures[0]["data"].shape == (8, 8, 4) and textures[1]["data"].dtype == np.uint8
    assert list(at) == [0, 1, -1, -1]
    m = scenes.cornell_mirror()
    assert (m["material"] == 4).sum() == 12 and list(m["material_type"]) == [0, 0, 0, 0, 1]
    assert np.array_equal(m["positions"], c["positions"])
