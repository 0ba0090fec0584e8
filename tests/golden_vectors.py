"""Golden vectors copied from the reference's own tests (values only; citations are into /root/reference)."""

# seqhasher_test.go:313-329 — TestGetHashFunc, input "ACTG"
ACTG = {
    "sha1": "65c89f59d38cdbf90dfaf0b0a6884829df8396b0",
    "sha3": "01eb915e4d8b6d44d0432c12dfdb949c1da1f37c295a653b8761a1e46ed2d76cb0c297d612af809b9691d341cad536df"
            "912cbba6e95a93380cdc9f545d9bfdcc",
    "md5": "86bfb9f78dd8b6cd35962bb7324fdbf8",
    "xxhash": "704b34bf20faedf2",
    "xxh3": "623952c8b43f0072",
    "cityhash": "7ee08b0605f909cf400644ddb3b8b80b",
    "murmur3": "da48f168029d0eff17c81eff7624a72f",
    "nthash": "508876b331232519",
    "blake3": "fe31e49d18b8883e7167198f770b98bba33b533cc12a9bb63ab264e5b70a347a",
    "k12": "9985653b7671edf4d1194c22886e3d1324c5ab28768c83e684387aea98c39ca0bceef0014d42ef5a3021b0686c537752"
           "9796e36401de254c977e14d59ca9d64ae567d568cd0f63f022a37b9678ac329d874860182f9e20fef9adeb0d3057c93b"
           "4f36386234798acfb94f469bee78baac64911e5ae1294db2b86aea32c5b3be12",
    "rapidhash": "e5ba56a5fa8a1dcf",
    "rapidhash32": "1f304b6a",
}

# (algo, input bytes, expected hex, citation)
EXTRA = [
    ("sha1", b"TGCA", "e3da52abc8fbdb38b113a187ed0ac763fa86d1d4", "seqhasher_test.go:265"),
    ("sha1", b"AAAA", "e2512172abf8cc9f67fdd49eb6cacf2df71bbad3", "test/test2_binary.sh:12"),
    ("sha1", b"aaaa", "70c881d4a26984ddce795f6f71817c9cf4480e79", "test/test2_binary.sh:97"),
    ("sha1", b"ACTGINVALID", "3e06752b", "seqhasher_test.go:834 (prefix)"),
    ("sha1", b"ACTG" * 25000, "3200abf5034212fb5d90fa242fc1cea5e0f3dd00", "test/test2_binary.sh:155"),
    ("sha1", b"A" * 100000, "e27b41eb7f7370b36f1b2fa63f05652f52423277", "test/test2_binary.sh:155"),
    ("md5", b"TGCA", "5c15f97a88433c48f8bf76745d9da437", "seqhasher_test.go:278"),
    ("xxhash", b"AAAA", "cf40b5b72bc43e77", "test/test2_binary.sh:80"),
    ("xxhash", b"aaaa", "42a70d1abf84bf32", "test/test2_binary.sh:80"),
    ("nthash", b"TGCA", "95cecc5106c8fccd", "seqhasher_test.go:290"),
]

# Published known-answer tests of the un-vendored dependencies (not in the reference repo)
MURMUR3_KATS = {  # MurmurHash3_x64_128 seed 0, printed h1||h2
    b"hello": "cbd8a7b341bd9b025b1e906a48ae1d19",
    b"hello, world": "342fac623a5ebc8e4cdcbc079642414d",
    b"19 Jan 2038 at 3:14:07 AM": "b89e5988b737affc664fc2950231b2cb",
    b"The quick brown fox jumps over the lazy dog.": "cd99481f9ee902c9695da1a38987b6e7",
    b"The quick brown fox jumps over the lazy dog": "e34bbc7bbc071b6c7a433ca9c49a9347",
}
# RFC 9861 KT128 test vectors (32-byte outputs; the reference reads 128, whose first 32 bytes are these)
def ptn(n: int) -> bytes:
    return bytes(i % 251 for i in range(n))
KT128_KATS = [
    (ptn(1), "2bda92450e8b147f8a7cb629e784a058efca7cf7d8218e02d345dfaa65244a1f"),
    (ptn(17), "6bf75fa2239198db4772e36478f8e19b0f371205f6a9a93a273f51df37122888"),
    (ptn(17 ** 2), "0c315ebcdedbf61426de7dcf8fb725d1e74675d7f5327a5067f367b108ecb67c"),
    (ptn(17 ** 3), "cb552e2ec77d9910701d578b457ddf772c12e322e4ee7fe417f92c758f0d59d0"),
    (ptn(17 ** 4), "8701045e22205345ff4dda05555cbb5c3af1a771c2b89baef37db43d9998b9fe"),
]
CITY128_ABC = "a085f09013029e453980b2afd2126c04"   # CityHash128("abc") v1.1, printed High||Low
CITY64_ABC = 2640714258260161385                    # CityHash64("abc")
RAPIDHASH_V1_HELLO_WORLD = 17498481775468162579     # upstream rapidhash v1 KAT (shared structure check)
