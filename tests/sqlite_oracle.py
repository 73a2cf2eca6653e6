"""SQLite-level oracle: the reference's schema and the verbatim similarity SQL, executed by
Python's sqlite3 with the C oracle registered as the `cosine_distance` scalar function.

Pins what upstream leaves to SQLite: WHERE dist < ?, ORDER BY dist ASC, LIMIT, tie order
(src/engine.rs:375-383), the table shapes (src/engine.rs:31-48) and the UDF contract (:608-622).
"""
import sqlite3

from oracle import oracle

IMAGE_SCHEMA_V1 = """CREATE TABLE images (
	id               INTEGER PRIMARY KEY,
	filename         TEXT NOT NULL,
	path             TEXT NOT NULL,
	image_width      INTEGER,
	image_height     INTEGER,
	thumbnail        BLOB,
	created          DATETIME,
	indexed          DATETIME,
	UNIQUE(path)
)"""
HASH_TABLE_SCHEMA_V1 = "CREATE TABLE semantic_hashes (image_id INTEGER PRIMARY KEY, hash BLOB)"
SELECT_FIELDS = """
	images.id,
	images.filename,
	images.path,
	images.image_width,
	images.image_height,
	images.thumbnail
"""


def similarity_sql(limit: int = 100) -> str:
    # src/engine.rs:375-382; upstream hard-codes LIMIT 100, tests substitute k.
    return f"""
			SELECT {SELECT_FIELDS}, semantic_hashes.hash, cosine_distance(?, semantic_hashes.hash) AS dist
			FROM semantic_hashes
			INNER JOIN images images ON images.id = semantic_hashes.image_id
			WHERE dist < ?
			ORDER BY dist ASC
			LIMIT {int(limit)}"""


def make_db(path: str, ids, hashes) -> sqlite3.Connection:
    """Create a DB with the reference schema and one image + one hash row per entry."""
    conn = sqlite3.connect(path)
    conn.execute(IMAGE_SCHEMA_V1)
    conn.execute(HASH_TABLE_SCHEMA_V1)
    conn.executemany(
        "INSERT INTO images (id, filename, path, image_width, image_height, thumbnail) VALUES (?, ?, ?, 256, 256, x'00')",
        [(int(i), f"img{int(i)}.png", f"/synthetic/img{int(i)}.png") for i in ids])
    conn.executemany("INSERT OR IGNORE INTO semantic_hashes (image_id, hash) VALUES (?, ?)",
                     [(int(i), bytes(h)) for i, h in zip(ids, hashes)])
    conn.commit()
    register(conn)
    return conn


def register(conn: sqlite3.Connection) -> None:
    conn.create_function("cosine_distance", 2, oracle.udf_cosine_distance, deterministic=True)


def query(conn: sqlite3.Connection, query_hash: bytes, max_dist: float = 1e3, limit: int = 100):
    """Returns [(image_id, dist_f64)] exactly as engine.rs:383-387 would see the rows."""
    rows = conn.execute(similarity_sql(limit), (bytes(query_hash), float(max_dist))).fetchall()
    return [(r[0], r[7]) for r in rows]
