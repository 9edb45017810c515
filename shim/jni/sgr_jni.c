/* sgr_jni.c — JNI glue between surge.gpu.Native (shim/scala) and include/sgr.h.
 * Built only where a JDK exists (needs <jni.h>):  gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux
 *     -I../../include sgr_jni.c -L../../surge_b200/lib -lsgr -o libsgr_jni.so
 * The build image of this repository has no JDK, so this file is guarded and compiled nowhere here; all logic lives
 * behind the C ABI, which the Python/ctypes tests cover. */
#if defined(__has_include)
#if __has_include(<jni.h>)
#include <jni.h>
#include <stdint.h>
#include <string.h>

#include "sgr.h"

#define H(h) ((sgr_engine*)(intptr_t)(h))

/* Lengths and buffers come from Java: check them here, throw IllegalArgumentException, never read past what was handed over
 * (round-1 review: a short `firsts` array, a non-direct buffer (GetDirectBufferAddress == NULL) or an nbytes beyond the
 * buffer's capacity went straight to the C ABI). */
static int bad_arg(JNIEnv* env, const char* what) {
  (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/IllegalArgumentException"), what);
  return SGR_ERR_INVALID;
}
/* direct buffer address, checked: non-NULL and at least `need` bytes of capacity */
static void* direct(JNIEnv* env, jobject buf, jlong need, const char* what, int* ok) {
  void* p = buf ? (*env)->GetDirectBufferAddress(env, buf) : 0;
  if (!p || need < 0 || (*env)->GetDirectBufferCapacity(env, buf) < need) { bad_arg(env, what); *ok = 0; return 0; }
  return p;
}

static void throw_for(JNIEnv* env, sgr_engine* e, int32_t rc) {
  const char* cls = rc == SGR_ERR_STATE ? "org/apache/kafka/streams/errors/InvalidStateStoreException" : "java/lang/RuntimeException";
  (*env)->ThrowNew(env, (*env)->FindClass(env, cls), sgr_last_error(e));
}

JNIEXPORT jlong JNICALL Java_surge_gpu_Native_00024_create(JNIEnv* env, jobject o, jint device) {
  sgr_config cfg; memset(&cfg, 0, sizeof cfg); cfg.device = device;
  sgr_engine* e = 0;
  int32_t rc = sgr_create(&cfg, &e);
  if (rc != SGR_OK) { throw_for(env, 0, rc); return 0; }  /* SGR_ERR_NO_DEVICE: fail loudly, never fall back */
  return (jlong)(intptr_t)e;
}
JNIEXPORT jint JNICALL Java_surge_gpu_Native_00024_destroy(JNIEnv* env, jobject o, jlong h) { return sgr_destroy(H(h)); }
JNIEXPORT jstring JNICALL Java_surge_gpu_Native_00024_lastError(JNIEnv* env, jobject o, jlong h) { return (*env)->NewStringUTF(env, sgr_last_error(H(h))); }
JNIEXPORT jint JNICALL Java_surge_gpu_Native_00024_registerProgram(JNIEnv* env, jobject o, jlong h, jobject prog) {
  int ok = 1; void* p = direct(env, prog, (jlong)sizeof(sgr_fold_program), "program: direct buffer of sizeof(sgr_fold_program) bytes", &ok);
  return ok ? sgr_register_program(H(h), (const sgr_fold_program*)p) : SGR_ERR_INVALID;
}
JNIEXPORT jint JNICALL Java_surge_gpu_Native_00024_loadEvents(JNIEnv* env, jobject o, jlong h, jobject ev, jlong nbytes, jobject offs, jlong n_agg) {
  int ok = 1;
  void* e = direct(env, ev, nbytes, "events: direct buffer shorter than nbytes", &ok);
  void* so = ok && n_agg >= 0 ? direct(env, offs, (n_agg + 1) * 8, "segOffsets: direct buffer of (nAgg + 1) u64", &ok) : 0;
  return ok && so ? sgr_load_events(H(h), e, (uint64_t)nbytes, (const uint64_t*)so, (uint64_t)n_agg) : SGR_ERR_INVALID;
}
JNIEXPORT jint JNICALL Java_surge_gpu_Native_00024_loadUnsorted(JNIEnv* env, jobject o, jlong h, jobject rec, jlong n, jlong n_agg) {
  int ok = 1; void* r = n >= 0 && n <= (INT64_MAX / 64) ? direct(env, rec, n * 64, "records: direct buffer shorter than nRecords * 64", &ok) : (bad_arg(env, "nRecords"), (void*)0);
  return ok && r ? sgr_load_unsorted(H(h), r, (uint64_t)n, (uint64_t)n_agg) : SGR_ERR_INVALID;
}
JNIEXPORT jint JNICALL Java_surge_gpu_Native_00024_setInitialStates(JNIEnv* env, jobject o, jlong h, jobject st, jlong n_agg) {
  if (!st) return sgr_set_initial_states(H(h), 0, 0);
  int ok = 1; void* p = direct(env, st, 16 * n_agg, "states: direct buffer shorter than nAgg states", &ok);   /* >= 16 bytes per state */
  return ok ? sgr_set_initial_states(H(h), p, (uint64_t)n_agg) : SGR_ERR_INVALID;
}
JNIEXPORT jint JNICALL Java_surge_gpu_Native_00024_fold(JNIEnv* env, jobject o, jlong h) { return sgr_fold(H(h)); }
JNIEXPORT jint JNICALL Java_surge_gpu_Native_00024_foldIncremental(JNIEnv* env, jobject o, jlong h, jobject rec, jlong n) {
  int ok = 1; void* r = n >= 0 && n <= (INT64_MAX / 64) ? direct(env, rec, n * 64, "records: direct buffer shorter than nRecords * 64", &ok) : (bad_arg(env, "nRecords"), (void*)0);
  return ok && r ? sgr_fold_incremental(H(h), r, (uint64_t)n) : SGR_ERR_INVALID;
}
JNIEXPORT jint JNICALL Java_surge_gpu_Native_00024_loadKeys(JNIEnv* env, jobject o, jlong h, jobject keys, jobject offs, jlong n) {
  int ok = 1;
  const uint32_t* ko = n >= 0 ? (const uint32_t*)direct(env, offs, (n + 1) * 4, "keyOffsets: direct buffer of (n + 1) u32", &ok) : 0;
  if (!ok || !ko) return SGR_ERR_INVALID;
  const uint8_t* k = (const uint8_t*)direct(env, keys, (jlong)ko[n], "keys: direct buffer shorter than keyOffsets[n]", &ok);
  return ok ? sgr_load_keys(H(h), k, ko, (uint64_t)n) : SGR_ERR_INVALID;
}
JNIEXPORT jbyteArray JNICALL Java_surge_gpu_Native_00024_get(JNIEnv* env, jobject o, jlong h, jbyteArray key) {
  jsize klen = (*env)->GetArrayLength(env, key);
  jbyte* k = (*env)->GetByteArrayElements(env, key, 0);
  uint8_t out[SGR_MAX_STATE_BYTES]; uint32_t outlen = 0; int32_t exists = 0;
  int32_t rc = sgr_get(H(h), (const uint8_t*)k, (uint32_t)klen, out, sizeof out, &outlen, &exists);
  (*env)->ReleaseByteArrayElements(env, key, k, JNI_ABORT);
  if (rc != SGR_OK) { throw_for(env, H(h), rc); return 0; }
  if (!exists) return 0;                                   /* Option.empty */
  jbyteArray r = (*env)->NewByteArray(env, (jsize)outlen);  /* a fresh array, as RocksDB returns */
  (*env)->SetByteArrayRegion(env, r, 0, (jsize)outlen, (const jbyte*)out);
  return r;
}
JNIEXPORT jint JNICALL Java_surge_gpu_Native_00024_exportStates(JNIEnv* env, jobject o, jlong h, jobject out, jobject changed) {
  return sgr_export_states(H(h), (*env)->GetDirectBufferAddress(env, out), (uint64_t)(*env)->GetDirectBufferCapacity(env, out), 0,
                           changed ? (uint8_t*)(*env)->GetDirectBufferAddress(env, changed) : 0, 0);
}
JNIEXPORT jint JNICALL Java_surge_gpu_Native_00024_partitionForKey(JNIEnv* env, jobject o, jbyteArray key, jint n, jboolean up_to_colon) {
  jsize klen = (*env)->GetArrayLength(env, key);
  jbyte* k = (*env)->GetByteArrayElements(env, key, 0);
  int32_t p = -1;
  int32_t rc = n > 0 ? sgr_partition_for_key_utf8((const uint8_t*)k, (uint32_t)klen, (uint32_t)n, up_to_colon ? 1 : 0, &p) : SGR_ERR_INVALID;
  (*env)->ReleaseByteArrayElements(env, key, k, JNI_ABORT);
  if (rc != SGR_OK) { bad_arg(env, "partitionForKey: numPartitions must be positive and the key valid UTF-8"); return -1; }
  return p;
}

/* ---- ingest (raw Kafka record batches) */
#define G(g) ((sgr_ingest*)(intptr_t)(g))
JNIEXPORT jlong JNICALL Java_surge_gpu_Native_00024_ingestCreate(JNIEnv* env, jobject o) {
  sgr_ingest* g = 0;
  if (sgr_ingest_create(&g) != SGR_OK) { (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/OutOfMemoryError"), "sgr_ingest_create"); return 0; }
  return (jlong)(intptr_t)g;
}
JNIEXPORT jint JNICALL Java_surge_gpu_Native_00024_ingestDestroy(JNIEnv* env, jobject o, jlong g) { return sgr_ingest_destroy(G(g)); }
JNIEXPORT jint JNICALL Java_surge_gpu_Native_00024_ingestSetValueFraming(JNIEnv* env, jobject o, jlong g, jint framing) { return sgr_ingest_set_value_framing(G(g), framing); }
JNIEXPORT jint JNICALL Java_surge_gpu_Native_00024_ingestSetNullValueType(JNIEnv* env, jobject o, jlong g, jint event_type) { return sgr_ingest_set_null_value_type(G(g), event_type); }
/* sgr_ingest_set_json_packer takes an array of structs with strings: bind it with a small marshaller (or JNA) on the maintainer's side. */
JNIEXPORT jint JNICALL Java_surge_gpu_Native_00024_ingestSetAborted(JNIEnv* env, jobject o, jlong g, jint partition, jlongArray pids, jlongArray firsts) {
  jsize n = (*env)->GetArrayLength(env, pids);
  if ((*env)->GetArrayLength(env, firsts) != n) return bad_arg(env, "producerIds and firstOffsets differ in length");
  jlong* p = (*env)->GetLongArrayElements(env, pids, 0);
  jlong* f = (*env)->GetLongArrayElements(env, firsts, 0);
  int32_t rc = sgr_ingest_set_aborted(G(g), partition, (const int64_t*)p, (const int64_t*)f, (uint64_t)n);
  (*env)->ReleaseLongArrayElements(env, pids, p, JNI_ABORT);
  (*env)->ReleaseLongArrayElements(env, firsts, f, JNI_ABORT);
  return rc;
}
JNIEXPORT jlong JNICALL Java_surge_gpu_Native_00024_ingestRecordBatches(JNIEnv* env, jobject o, jlong g, jint partition, jobject data, jlong nbytes) {
  sgr_ingest_stats st;
  int ok = 1; void* d = direct(env, data, nbytes, "data: direct buffer shorter than nbytes", &ok);
  if (!ok) return -1;
  int32_t rc = sgr_ingest_record_batches(G(g), partition, d, (uint64_t)nbytes, &st);
  if (rc != SGR_OK) {   /* a corrupt batch kills the stream thread, as a CorruptRecordException would */
    (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/RuntimeException"), sgr_ingest_last_error(G(g)));
    return -1;
  }
  return (jlong)st.n_records;
}
JNIEXPORT jint JNICALL Java_surge_gpu_Native_00024_foldIngested(JNIEnv* env, jobject o, jlong h, jlong g) { return sgr_fold_ingested(H(h), G(g)); }
JNIEXPORT jint JNICALL Java_surge_gpu_Native_00024_growStates(JNIEnv* env, jobject o, jlong h, jlong n_agg) { return sgr_grow_states(H(h), (uint64_t)n_agg); }
JNIEXPORT jlongArray JNICALL Java_surge_gpu_Native_00024_ingestOffsets(JNIEnv* env, jobject o, jlong g, jint partition) {
  int64_t v[2] = {0, 0};
  sgr_ingest_offsets(G(g), partition, &v[0], &v[1]);
  jlongArray r = (*env)->NewLongArray(env, 2);
  (*env)->SetLongArrayRegion(env, r, 0, 2, (const jlong*)v);
  return r;
}

/* ---- device ingest: the same bytes, decoded on the GPU (include/sgr.h "device ingest") */
#define DG(g) ((sgr_dingest*)(intptr_t)(g))
JNIEXPORT jlong JNICALL Java_surge_gpu_Native_00024_dingestCreate(JNIEnv* env, jobject o, jlong h, jlong max_keys, jlong max_id_bytes) {
  sgr_dingest* g = 0;
  if (max_keys <= 0 || max_id_bytes < 0) { bad_arg(env, "maxKeys must be positive, maxIdBytes non-negative"); return 0; }
  int32_t rc = sgr_dingest_create(H(h), (uint64_t)max_keys, (uint64_t)max_id_bytes, &g);
  if (rc != SGR_OK) { (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/RuntimeException"), "sgr_dingest_create failed (no device memory, or no engine)"); return 0; }
  return (jlong)(intptr_t)g;
}
JNIEXPORT jint JNICALL Java_surge_gpu_Native_00024_dingestDestroy(JNIEnv* env, jobject o, jlong g) { return sgr_dingest_destroy(DG(g)); }
JNIEXPORT jint JNICALL Java_surge_gpu_Native_00024_dingestSetNullValueType(JNIEnv* env, jobject o, jlong g, jint event_type) { return sgr_dingest_set_null_value_type(DG(g), event_type); }
JNIEXPORT jint JNICALL Java_surge_gpu_Native_00024_dingestSetAborted(JNIEnv* env, jobject o, jlong g, jint partition, jlongArray pids, jlongArray firsts) {
  jsize n = (*env)->GetArrayLength(env, pids);
  if ((*env)->GetArrayLength(env, firsts) != n) return bad_arg(env, "producerIds and firstOffsets differ in length");
  jlong* p = (*env)->GetLongArrayElements(env, pids, 0);
  jlong* f = (*env)->GetLongArrayElements(env, firsts, 0);
  int32_t rc = sgr_dingest_set_aborted(DG(g), partition, (const int64_t*)p, (const int64_t*)f, (uint64_t)n);
  (*env)->ReleaseLongArrayElements(env, pids, p, JNI_ABORT);
  (*env)->ReleaseLongArrayElements(env, firsts, f, JNI_ABORT);
  return rc;
}
/* `data` must be a DIRECT buffer that stays untouched until dingestFold returns: the copy to the device is asynchronous.
 * Returns the number of data batches queued; throws on a malformed fetch (nothing of it is queued). */
JNIEXPORT jlong JNICALL Java_surge_gpu_Native_00024_dingestSubmit(JNIEnv* env, jobject o, jlong g, jint partition, jobject data, jlong nbytes) {
  sgr_ingest_stats st;
  int ok = 1; void* d = direct(env, data, nbytes, "data: direct buffer shorter than nbytes", &ok);
  if (!ok) return -1;
  int32_t rc = sgr_dingest_submit(DG(g), partition, d, (uint64_t)nbytes, &st);
  if (rc != SGR_OK) {
    (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/RuntimeException"), sgr_dingest_last_error(DG(g)));
    return -1;
  }
  return (jlong)st.n_batches;
}
/* decode + intern + fold of everything submitted; Array(records folded, new aggregate ids). A corrupt batch fails the whole poll
 * (nothing applied, positions unchanged) and kills the consuming thread, as a CorruptRecordException would. */
JNIEXPORT jlongArray JNICALL Java_surge_gpu_Native_00024_dingestFold(JNIEnv* env, jobject o, jlong g) {
  sgr_ingest_stats st;
  int32_t rc = sgr_dingest_fold(DG(g), &st);
  if (rc != SGR_OK) {
    (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/RuntimeException"), sgr_dingest_last_error(DG(g)));
    return 0;
  }
  int64_t v[2] = {(int64_t)st.n_records, (int64_t)st.n_new_keys};
  jlongArray r = (*env)->NewLongArray(env, 2);
  (*env)->SetLongArrayRegion(env, r, 0, 2, (const jlong*)v);
  return r;
}
JNIEXPORT jlongArray JNICALL Java_surge_gpu_Native_00024_dingestOffsets(JNIEnv* env, jobject o, jlong g, jint partition) {
  int64_t v[2] = {0, 0};
  sgr_dingest_offsets(DG(g), partition, &v[0], &v[1]);
  jlongArray r = (*env)->NewLongArray(env, 2);
  (*env)->SetLongArrayRegion(env, r, 0, 2, (const jlong*)v);
  return r;
}
JNIEXPORT jint JNICALL Java_surge_gpu_Native_00024_dingestReset(JNIEnv* env, jobject o, jlong g) { return sgr_dingest_reset(DG(g)); }
#endif
#endif
