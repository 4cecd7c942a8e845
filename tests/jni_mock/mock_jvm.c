/*
 * mock_jvm.c — a JNIEnv over plain C arrays + the call sequence of B200LambdaMART.init()/learn()
 * (jni/java/ciir/umass/edu/learning/tree/B200LambdaMART.java), so that jni/ranklib_b200_jni.c is compiled and
 * executed by the test suite although this image has no JVM.  Test infrastructure only.
 *
 * Element-copy semantics are the strict ones a JVM may choose: Get*ArrayElements / GetPrimitiveArrayCritical hand
 * out a COPY, and Release* writes it back unless the mode is JNI_ABORT — a shim that forgets a release, releases
 * with the wrong mode, or touches an array after releasing it fails here (outstanding-pin counter, stale data).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "jni.h"

enum { OBJ_ARRAY = 1, OBJ_STRING, OBJ_CLASS, OBJ_THROWABLE };

struct mock_object {
    int kind;
    int elem;      /* bytes per element (arrays) */
    jsize len;     /* elements (arrays) */
    void* data;    /* array payload / string bytes */
    void* pinned;  /* outstanding copy handed to native code */
};
struct mock_method {
    char name[64];
};

static struct {
    int pins;            /* Get* without Release* */
    int pin_errors;      /* double pins, releases of unknown pointers */
    int thrown;          /* Throw calls */
    char message[1024];  /* message of the last thrown RankLibError */
    char klass[128];     /* class the shim looked up */
    char method[64], sig[160];
} J;

static struct mock_object the_class = {OBJ_CLASS, 0, 0, NULL, NULL};
static struct mock_method the_method;

static jclass m_FindClass(JNIEnv* env, const char* name) {
    (void)env;
    snprintf(J.klass, sizeof J.klass, "%s", name);
    return &the_class;
}
static jmethodID m_GetStaticMethodID(JNIEnv* env, jclass cls, const char* name, const char* sig) {
    (void)env; (void)cls;
    snprintf(J.method, sizeof J.method, "%s", name);
    snprintf(J.sig, sizeof J.sig, "%s", sig);
    return &the_method;
}
static jstring m_NewStringUTF(JNIEnv* env, const char* utf) {
    (void)env;
    struct mock_object* s = calloc(1, sizeof *s);
    s->kind = OBJ_STRING;
    s->data = strdup(utf ? utf : "");
    return s;
}
/* RankLibError.create(String) -> a throwable carrying the message */
static jobject m_CallStaticObjectMethod(JNIEnv* env, jclass cls, jmethodID m, ...) {
    (void)env; (void)cls; (void)m;
    va_list ap;
    va_start(ap, m);
    jstring msg = va_arg(ap, jstring);
    va_end(ap);
    struct mock_object* t = calloc(1, sizeof *t);
    t->kind = OBJ_THROWABLE;
    t->data = strdup(msg && msg->kind == OBJ_STRING ? (const char*)msg->data : "<not a string>");
    return t;
}
static jint m_Throw(JNIEnv* env, jthrowable obj) {
    (void)env;
    J.thrown++;
    snprintf(J.message, sizeof J.message, "%s", obj && obj->kind == OBJ_THROWABLE ? (const char*)obj->data : "<not a throwable>");
    return 0;
}
static jsize m_GetArrayLength(JNIEnv* env, jarray a) {
    (void)env;
    return a->len;
}
static void* pin(jarray a) {
    if (a->pinned) J.pin_errors++;
    size_t bytes = (size_t)a->len * (size_t)a->elem;
    a->pinned = malloc(bytes ? bytes : 1);
    memcpy(a->pinned, a->data, bytes);
    J.pins++;
    return a->pinned;
}
static void unpin(jarray a, void* p, jint mode) {
    if (!a->pinned || a->pinned != p) {
        J.pin_errors++;
        return;
    }
    if (mode != JNI_ABORT) memcpy(a->data, p, (size_t)a->len * (size_t)a->elem);
    if (mode != JNI_COMMIT) {
        memset(p, 0xA5, (size_t)a->len * (size_t)a->elem); /* a use after release reads garbage */
        free(p);
        a->pinned = NULL;
        J.pins--;
    }
}
static void* m_GetPrimitiveArrayCritical(JNIEnv* env, jarray a, jboolean* isCopy) {
    (void)env;
    if (isCopy) *isCopy = 1;
    return pin(a);
}
static void m_ReleasePrimitiveArrayCritical(JNIEnv* env, jarray a, void* p, jint mode) {
    (void)env;
    unpin(a, p, mode);
}
static jint* m_GetIntArrayElements(JNIEnv* env, jintArray a, jboolean* isCopy) { return m_GetPrimitiveArrayCritical(env, a, isCopy); }
static void m_ReleaseIntArrayElements(JNIEnv* env, jintArray a, jint* p, jint mode) { (void)env; unpin(a, p, mode); }
static jfloat* m_GetFloatArrayElements(JNIEnv* env, jfloatArray a, jboolean* isCopy) { return m_GetPrimitiveArrayCritical(env, a, isCopy); }
static void m_ReleaseFloatArrayElements(JNIEnv* env, jfloatArray a, jfloat* p, jint mode) { (void)env; unpin(a, p, mode); }
static void m_SetIntArrayRegion(JNIEnv* env, jintArray a, jsize start, jsize len, const jint* buf) {
    (void)env;
    if (start < 0 || len < 0 || start + len > a->len) {
        J.pin_errors++;
        return;
    }
    memcpy((jint*)a->data + start, buf, (size_t)len * sizeof(jint));
}

static void m_SetByteArrayRegion(JNIEnv* env, jbyteArray a, jsize start, jsize len, const jbyte* buf) {
    (void)env;
    if (start < 0 || len < 0 || start + len > a->len || a->elem != 1) {
        J.pin_errors++;
        return;
    }
    memcpy((jbyte*)a->data + start, buf, (size_t)len);
}
static void m_GetByteArrayRegion(JNIEnv* env, jbyteArray a, jsize start, jsize len, jbyte* buf) {
    (void)env;
    if (start < 0 || len < 0 || start + len > a->len || a->elem != 1) {
        J.pin_errors++;
        return;
    }
    memcpy(buf, (jbyte*)a->data + start, (size_t)len);
}

static const char* m_GetStringUTFChars(JNIEnv* env, jstring s, jboolean* isCopy) {
    (void)env;
    if (isCopy) *isCopy = 1;
    if (!s || s->kind != OBJ_STRING || s->pinned) {
        J.pin_errors++;
        return NULL;
    }
    s->pinned = strdup((const char*)s->data);
    J.pins++;
    return (const char*)s->pinned;
}
static void m_ReleaseStringUTFChars(JNIEnv* env, jstring s, const char* utf) {
    (void)env;
    if (!s || s->pinned != (void*)utf) {
        J.pin_errors++;
        return;
    }
    free(s->pinned);
    s->pinned = NULL;
    J.pins--;
}

static const struct JNINativeInterface_ table = {
    m_FindClass, m_GetStaticMethodID, m_CallStaticObjectMethod, m_NewStringUTF, m_Throw, m_GetArrayLength,
    m_GetPrimitiveArrayCritical, m_ReleasePrimitiveArrayCritical, m_GetIntArrayElements, m_ReleaseIntArrayElements,
    m_GetFloatArrayElements, m_ReleaseFloatArrayElements, m_SetIntArrayRegion, m_GetStringUTFChars, m_ReleaseStringUTFChars,
    m_SetByteArrayRegion, m_GetByteArrayRegion};
static JNIEnv the_env = &table;

static jarray new_array(int elem, jsize len, const void* init) {
    struct mock_object* a = calloc(1, sizeof *a);
    a->kind = OBJ_ARRAY;
    a->elem = elem;
    a->len = len;
    a->data = calloc((size_t)(len > 0 ? len : 1), (size_t)elem);
    if (init) memcpy(a->data, init, (size_t)len * (size_t)elem);
    return a;
}
static void free_array(jarray a) {
    if (!a) return;
    free(a->pinned);
    free(a->data);
    free(a);
}

/* the natives of the shim (declared by NativeBridge.java) */
#define BRIDGE(name) Java_ciir_umass_edu_learning_tree_NativeBridge_##name
jlong BRIDGE(create)(JNIEnv*, jclass, jint);
jint BRIDGE(destroy)(JNIEnv*, jclass, jlong);
jint BRIDGE(loadDense)(JNIEnv*, jclass, jlong, jfloatArray, jlong, jint, jintArray, jfloatArray, jintArray);
jint BRIDGE(init)(JNIEnv*, jclass, jlong, jint, jint, jfloat, jint, jint, jint, jint, jfloat, jlong);
jint BRIDGE(loadLetorFile)(JNIEnv*, jclass, jlong, jstring, jboolean, jintArray, jintArray);
jfloat BRIDGE(boostIter)(JNIEnv*, jclass, jlong, jintArray, jfloatArray, jdoubleArray, jintArray);
jint BRIDGE(readScores)(JNIEnv*, jclass, jlong, jdoubleArray);
jint BRIDGE(ensembleEval)(JNIEnv*, jclass, jlong, jintArray, jfloatArray, jintArray, jfloatArray, jfloatArray, jlong, jint, jfloatArray);
jint BRIDGE(loadValidation)(JNIEnv*, jclass, jlong, jfloatArray, jlong, jint, jfloatArray, jintArray);
jfloat BRIDGE(validMetric)(JNIEnv*, jclass, jlong);
jdouble BRIDGE(scoreResident)(JNIEnv*, jclass, jlong, jint, jintArray, jfloatArray, jintArray, jfloatArray);
jint BRIDGE(loadBag)(JNIEnv*, jclass, jlong, jlong, jintArray);
jint BRIDGE(commUniqueId)(JNIEnv*, jclass, jbyteArray);
jint BRIDGE(commInit)(JNIEnv*, jclass, jlong, jint, jint, jbyteArray);

/* --- what the Python test reads back --- */
int mock_thrown(void) { return J.thrown; }
const char* mock_message(void) { return J.message; }
const char* mock_error_class(void) { return J.klass; }
const char* mock_error_method(void) { return J.method; }
const char* mock_error_sig(void) { return J.sig; }
int mock_outstanding_pins(void) { return J.pins; }
int mock_pin_errors(void) { return J.pin_errors; }
void mock_reset(void) { memset(&J, 0, sizeof J); }

/*
 * B200LambdaMART.init() + learn() for n_trees trees (no validation set), then readScores, then one batched
 * Ensemble.eval of the training rows with the trees just learnt — every native of NativeBridge once.
 *   X [N][F] row-major (column j = fid j+1), labels [N], qoff [Q+1]
 *   out: node_ints [n_trees][cap*7], node_floats [n_trees][cap*2], node_dev [n_trees][cap], n_nodes [n_trees],
 *        metric [n_trees], scores [N], eval_out [N]
 * Returns 0, or the ordinal (1..) of the native call after which a Java exception was pending.
 */
int mock_train(int device, const float* X, long long N, int F, const float* labels, const int* qoff, int Q, int n_leaves,
               int min_leaf_support, float lr, int n_threshold, int kind, int metric, int k, int n_trees, int* node_ints,
               float* node_floats, double* node_dev, int* n_nodes, float* metric_out, double* scores, float* eval_out) {
    JNIEnv* env = &the_env;
    int step = 0, rc = 0;
    int cap = 2 * n_leaves + 1; /* NativeBridge.nodeCapacity */
    jlong h = 0;
    jarray jx = NULL, jf = NULL, jl = NULL, jq = NULL, jni_ = NULL, jnf = NULL, jnd = NULL, jnn = NULL, jsc = NULL;
    jarray ani = NULL, anf = NULL, ato = NULL, aw = NULL, ax = NULL, ao = NULL;
    int* fids = malloc(sizeof(int) * (size_t)F);
    for (int j = 0; j < F; j++) fids[j] = j + 1;

    step++;
    h = BRIDGE(create)(env, NULL, device);
    if (J.thrown) { rc = step; goto done; }

    jx = new_array(4, (jsize)(N * F), X);
    jf = new_array(4, F, fids);
    jl = new_array(4, (jsize)N, labels);
    jq = new_array(4, Q + 1, qoff);
    step++;
    BRIDGE(loadDense)(env, NULL, h, jx, N, F, jf, jl, jq);
    if (J.thrown) { rc = step; goto done; }

    step++;
    BRIDGE(init)(env, NULL, h, n_leaves, min_leaf_support, lr, n_threshold, kind, metric, k, 1.0f, 0);
    if (J.thrown) { rc = step; goto done; }

    jni_ = new_array(4, cap * 7, NULL);
    jnf = new_array(4, cap * 2, NULL);
    jnd = new_array(8, cap, NULL);
    jnn = new_array(4, 1, NULL);
    for (int t = 0; t < n_trees; t++) {
        step++;
        metric_out[t] = BRIDGE(boostIter)(env, NULL, h, jni_, jnf, jnd, jnn);
        if (J.thrown) { rc = step; goto done; }
        n_nodes[t] = ((int*)jnn->data)[0];
        memcpy(node_ints + (size_t)t * cap * 7, jni_->data, sizeof(int) * (size_t)cap * 7);
        memcpy(node_floats + (size_t)t * cap * 2, jnf->data, sizeof(float) * (size_t)cap * 2);
        memcpy(node_dev + (size_t)t * cap, jnd->data, sizeof(double) * (size_t)cap);
    }

    jsc = new_array(8, (jsize)N, NULL);
    step++;
    BRIDGE(readScores)(env, NULL, h, jsc);
    if (J.thrown) { rc = step; goto done; }
    memcpy(scores, jsc->data, sizeof(double) * (size_t)N);

    /* Ensemble.eval over the learnt trees: concatenate the node arrays, weights = learningRate each; the feature
     * matrix is indexed by fid directly (column 0 unused) */
    {
        int total = 0;
        for (int t = 0; t < n_trees; t++) total += n_nodes[t];
        int* ci = calloc((size_t)total * 7 + 1, sizeof(int));
        float* cf = calloc((size_t)total * 2 + 1, sizeof(float));
        int* off = calloc((size_t)n_trees + 1, sizeof(int));
        float* w = calloc((size_t)n_trees + 1, sizeof(float));
        float* xe = calloc((size_t)N * (size_t)(F + 1), sizeof(float));
        int at = 0;
        for (int t = 0; t < n_trees; t++) {
            off[t] = at;
            w[t] = lr;
            memcpy(ci + (size_t)at * 7, node_ints + (size_t)t * cap * 7, sizeof(int) * (size_t)n_nodes[t] * 7);
            memcpy(cf + (size_t)at * 2, node_floats + (size_t)t * cap * 2, sizeof(float) * (size_t)n_nodes[t] * 2);
            at += n_nodes[t];
        }
        off[n_trees] = at;
        for (long long i = 0; i < N; i++) memcpy(xe + i * (F + 1) + 1, X + i * F, sizeof(float) * (size_t)F);
        ani = new_array(4, total * 7, ci);
        anf = new_array(4, total * 2, cf);
        ato = new_array(4, n_trees + 1, off);
        aw = new_array(4, n_trees, w);
        ax = new_array(4, (jsize)(N * (F + 1)), xe);
        ao = new_array(4, (jsize)N, NULL);
        free(ci); free(cf); free(off); free(w); free(xe);
        step++;
        BRIDGE(ensembleEval)(env, NULL, h, ani, anf, ato, aw, ax, N, F + 1, ao);
        if (J.thrown) { rc = step; goto done; }
        memcpy(eval_out, ao->data, sizeof(float) * (size_t)N);
    }

done:
    if (h) BRIDGE(destroy)(env, NULL, h);
    free(fids);
    free_array(jx); free_array(jf); free_array(jl); free_array(jq);
    free_array(jni_); free_array(jnf); free_array(jnd); free_array(jnn); free_array(jsc);
    free_array(ani); free_array(anf); free_array(ato); free_array(aw); free_array(ax); free_array(ao);
    return rc;
}

/*
 * B200LambdaMART.initFromFile(path) + n_trees x boostIter: the file goes to the device through NativeBridge.loadLetorFile.
 *   out: dims[3] = N, Q, maxFid; node_ints / node_floats / n_nodes / metric_out as in mock_train.
 * Returns 0, or the ordinal (1..) of the native call after which a Java exception was pending.
 */
int mock_train_from_file(int device, const char* path, int must_have_rel, const int* features, int n_features, int n_leaves, int n_trees,
                         int* dims, int* node_ints, float* node_floats, int* n_nodes, float* metric_out) {
    JNIEnv* env = &the_env;
    int step = 0, rc = 0;
    int cap = 2 * n_leaves + 1;
    jlong h = 0;
    jarray jf = NULL, jd = NULL, jni_ = NULL, jnf = NULL, jnd = NULL, jnn = NULL;
    jstring jpath = m_NewStringUTF(env, path);
    step++;
    h = BRIDGE(create)(env, NULL, device);
    if (J.thrown) { rc = step; goto done; }
    if (features) jf = new_array(4, n_features, features);
    jd = new_array(4, 3, NULL);
    step++;
    BRIDGE(loadLetorFile)(env, NULL, h, jpath, (jboolean)must_have_rel, jf, jd);
    if (J.thrown) { rc = step; goto done; }
    memcpy(dims, jd->data, 3 * sizeof(int));
    step++;
    BRIDGE(init)(env, NULL, h, n_leaves, 1, 0.1f, 256, 0, 0, 10, 1.0f, 0);
    if (J.thrown) { rc = step; goto done; }
    jni_ = new_array(4, cap * 7, NULL);
    jnf = new_array(4, cap * 2, NULL);
    jnd = new_array(8, cap, NULL);
    jnn = new_array(4, 1, NULL);
    for (int t = 0; t < n_trees; t++) {
        step++;
        metric_out[t] = BRIDGE(boostIter)(env, NULL, h, jni_, jnf, jnd, jnn);
        if (J.thrown) { rc = step; goto done; }
        n_nodes[t] = ((int*)jnn->data)[0];
        memcpy(node_ints + (size_t)t * cap * 7, jni_->data, sizeof(int) * (size_t)cap * 7);
        memcpy(node_floats + (size_t)t * cap * 2, jnf->data, sizeof(float) * (size_t)cap * 2);
    }
done:
    if (h) BRIDGE(destroy)(env, NULL, h);
    free_array(jf); free_array(jd); free_array(jni_); free_array(jnf); free_array(jnd); free_array(jnn);
    if (jpath) { free(jpath->pinned); free(jpath->data); free(jpath); }
    return rc;
}

/*
 * B200LambdaMART.init() + learn() WITH a validation set, as jni/java/.../B200LambdaMART.java drives it: create, loadDense,
 * loadValidation, init, n_trees x (boostIter, validMetric), scoreResident(training), scoreResident(validation), destroy.
 * With picks != NULL the training context is a bag gathered on the device from a base context (B200RFRanker): create x 2,
 * loadDense(base), loadBag.
 *   out: node arrays as mock_train, valid_out[n_trees], final_scores[2] (training, validation)
 * Returns 0, or the ordinal of the native call after which a Java exception was pending.
 */
int mock_train_valid(int device, const float* X, long long N, int F, const float* labels, const int* qoff, int Q, const float* VX,
                     long long NV, const float* vlabels, const int* vqoff, int QV, const int* picks, int n_picks, int n_leaves,
                     int kind, int n_trees, int* node_ints, float* node_floats, int* n_nodes, float* metric_out, float* valid_out,
                     double* final_scores) {
    JNIEnv* env = &the_env;
    int step = 0, rc = 0;
    int cap = 2 * n_leaves + 1;
    jlong h = 0, base = 0;
    jarray jx = NULL, jf = NULL, jl = NULL, jq = NULL, jvx = NULL, jvl = NULL, jvq = NULL, jp = NULL;
    jarray jni_ = NULL, jnf = NULL, jnd = NULL, jnn = NULL, ani = NULL, anf = NULL, ato = NULL, aw = NULL;
    int* fids = malloc(sizeof(int) * (size_t)F);
    for (int j = 0; j < F; j++) fids[j] = j + 1;
    step++;
    h = BRIDGE(create)(env, NULL, device);
    if (J.thrown) { rc = step; goto done; }
    jx = new_array(4, (jsize)(N * F), X);
    jf = new_array(4, F, fids);
    jl = new_array(4, (jsize)N, labels);
    jq = new_array(4, Q + 1, qoff);
    if (picks) {
        step++;
        base = BRIDGE(create)(env, NULL, device);
        if (J.thrown) { rc = step; goto done; }
        step++;
        BRIDGE(loadDense)(env, NULL, base, jx, N, F, jf, jl, jq);
        if (J.thrown) { rc = step; goto done; }
        jp = new_array(4, n_picks, picks);
        step++;
        BRIDGE(loadBag)(env, NULL, h, base, jp);
        if (J.thrown) { rc = step; goto done; }
    } else {
        step++;
        BRIDGE(loadDense)(env, NULL, h, jx, N, F, jf, jl, jq);
        if (J.thrown) { rc = step; goto done; }
    }
    if (VX) {
        jvx = new_array(4, (jsize)(NV * F), VX);
        jvl = new_array(4, (jsize)NV, vlabels);
        jvq = new_array(4, QV + 1, vqoff);
        step++;
        BRIDGE(loadValidation)(env, NULL, h, jvx, NV, F, jvl, jvq);
        if (J.thrown) { rc = step; goto done; }
    }
    step++;
    BRIDGE(init)(env, NULL, h, n_leaves, 1, 0.1f, 256, kind, 0, 10, 1.0f, 0);
    if (J.thrown) { rc = step; goto done; }
    jni_ = new_array(4, cap * 7, NULL);
    jnf = new_array(4, cap * 2, NULL);
    jnd = new_array(8, cap, NULL);
    jnn = new_array(4, 1, NULL);
    for (int t = 0; t < n_trees; t++) {
        step++;
        metric_out[t] = BRIDGE(boostIter)(env, NULL, h, jni_, jnf, jnd, jnn);
        if (J.thrown) { rc = step; goto done; }
        n_nodes[t] = ((int*)jnn->data)[0];
        memcpy(node_ints + (size_t)t * cap * 7, jni_->data, sizeof(int) * (size_t)cap * 7);
        memcpy(node_floats + (size_t)t * cap * 2, jnf->data, sizeof(float) * (size_t)cap * 2);
        if (VX) {
            step++;
            valid_out[t] = BRIDGE(validMetric)(env, NULL, h);
            if (J.thrown) { rc = step; goto done; }
        }
    }
    {
        int total = 0;
        for (int t = 0; t < n_trees; t++) total += n_nodes[t];
        int* ci = calloc((size_t)total * 7 + 1, sizeof(int));
        float* cf = calloc((size_t)total * 2 + 1, sizeof(float));
        int* off = calloc((size_t)n_trees + 1, sizeof(int));
        float* w = calloc((size_t)n_trees + 1, sizeof(float));
        int at = 0;
        for (int t = 0; t < n_trees; t++) {
            off[t] = at;
            w[t] = 0.1f;
            memcpy(ci + (size_t)at * 7, node_ints + (size_t)t * cap * 7, sizeof(int) * (size_t)n_nodes[t] * 7);
            memcpy(cf + (size_t)at * 2, node_floats + (size_t)t * cap * 2, sizeof(float) * (size_t)n_nodes[t] * 2);
            at += n_nodes[t];
        }
        off[n_trees] = at;
        ani = new_array(4, total * 7, ci);
        anf = new_array(4, total * 2, cf);
        ato = new_array(4, n_trees + 1, off);
        aw = new_array(4, n_trees, w);
        free(ci); free(cf); free(off); free(w);
        step++;
        final_scores[0] = BRIDGE(scoreResident)(env, NULL, h, 0, ani, anf, ato, aw);
        if (J.thrown) { rc = step; goto done; }
        if (VX) {
            step++;
            final_scores[1] = BRIDGE(scoreResident)(env, NULL, h, 1, ani, anf, ato, aw);
            if (J.thrown) { rc = step; goto done; }
        }
    }
done:
    if (h) BRIDGE(destroy)(env, NULL, h);
    if (base) BRIDGE(destroy)(env, NULL, base);
    free(fids);
    free_array(jx); free_array(jf); free_array(jl); free_array(jq); free_array(jvx); free_array(jvl); free_array(jvq); free_array(jp);
    free_array(jni_); free_array(jnf); free_array(jnd); free_array(jnn); free_array(ani); free_array(anf); free_array(ato); free_array(aw);
    return rc;
}

/* NativeBridge.commUniqueId(byte[128]) then commInit(handle, 0, 1, id) on a fresh context (world size 1: no NCCL needed for
 * commInit; commUniqueId needs libnccl).  Also: boostIter with node arrays of inconsistent lengths must throw BEFORE the
 * library runs.  Returns 0 or the ordinal of the failing step; *id_nonzero = the id is not all zeros. */
int mock_comm_and_checks(int device, int* id_nonzero, int* bad_lengths_thrown) {
    JNIEnv* env = &the_env;
    int rc = 0, step = 0;
    jlong h = 0;
    jarray id = new_array(1, 128, NULL), ni = NULL, nf = NULL, nd = NULL, nn = NULL;
    step++;
    BRIDGE(commUniqueId)(env, NULL, id);
    if (J.thrown) { rc = step; goto done; }
    *id_nonzero = 0;
    for (int i = 0; i < 128; i++) *id_nonzero |= ((unsigned char*)id->data)[i] != 0;
    step++;
    h = BRIDGE(create)(env, NULL, device);
    if (J.thrown) { rc = step; goto done; }
    step++;
    BRIDGE(commInit)(env, NULL, h, 0, 1, id);
    if (J.thrown) { rc = step; goto done; }
    ni = new_array(4, 7 * 5, NULL);
    nf = new_array(4, 2 * 5, NULL);
    nd = new_array(8, 21, NULL);   /* 21 nodes of deviance but only 5 nodes of ints / floats */
    nn = new_array(4, 1, NULL);
    BRIDGE(boostIter)(env, NULL, h, ni, nf, nd, nn);
    *bad_lengths_thrown = J.thrown;
done:
    if (h) BRIDGE(destroy)(env, NULL, h);
    free_array(id); free_array(ni); free_array(nf); free_array(nd); free_array(nn);
    return rc;
}
