/*
 * ac_oracle.c — CPU ORACLE (test infrastructure, NOT a product path).  See ac_oracle.h.
 *
 * Literal C restatement of the reference's constructors and match loops, including
 * its node data structure (open-addressing FNV-1a HashmapNode, base+array RangeNode,
 * fail links, suffixMatch chain), so that it is also a faithful CPU baseline.
 * Citations are relative to /root/reference/src/main/java/com/roklenarcic/util/strings/.
 */
#include "ac_oracle.h"
#include "java_char_tables.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ JDK helpers */

static uint16_t g_lower[65536];
static int g_lower_ready = 0;

static void ensure_tables(void) {
    if (!g_lower_ready) {
        java_fill_lower_table(g_lower);
        g_lower_ready = 1;
    }
}

uint16_t ora_to_lower(uint16_t c) {
    ensure_tables();
    return g_lower[c];
}

int ora_is_letter_or_digit(uint16_t c) { return java_is_letter_or_digit(c); }

/* WordCharacters.java:6-16 (mode 0), :18-24 (mode 1), :26-39 (mode 2) */
void ora_word_chars(int mode, const uint16_t *chars, const uint8_t *toggles, int32_t n,
                    uint8_t *out) {
    memset(out, 0, 65536);
    if (mode == 0 || mode == 2) {
        out['-'] = 1;
        out['_'] = 1;
        for (int i = 0; i < 65536; i++) {
            if (java_is_letter_or_digit((uint16_t)i)) out[i] = 1;
        }
    }
    if (mode == 1) {
        for (int32_t i = 0; i < n; i++) out[chars[i]] = 1;
    } else if (mode == 2) {
        for (int32_t i = 0; i < n; i++) out[chars[i]] = toggles[i] ? 1 : 0;
    }
}

/* ------------------------------------------------------------------ nodes */

typedef struct Node Node;
struct Node {
    Node *defaultTransition; /* root ? this : null   (TrieNode ctor, AhoCorasickSet.java:505-507) */
    Node *failTransition;
    Node *suffixMatch;
    Node **children; /* HashmapNode.children / RangeNode.children */
    uint16_t *keys;  /* HashmapNode.keys */
    int32_t matchLength;
    int32_t level; /* LongestMatchSet.java:513 */
    int32_t value; /* index of the dictionary entry whose value object is stored; -1 = null/none */
    int32_t failMatchLength; /* WholeWordLongestMatchSet.java:TrieNode — last whole-word match up the tree */
    int32_t failMatchOffset;
    int32_t failValue;       /* WholeWordLongestMatchMap.java: value of that match */
    int32_t cap;   /* keys.length */
    int32_t modulusMask;
    int32_t numEntries;
    int32_t size; /* RangeNode.size */
    uint16_t baseChar;
    uint8_t kind; /* 0 = HashmapNode, 1 = RangeNode */
};

/* simple bump arena so that 10^6-keyword dictionaries build quickly */
typedef struct Chunk {
    struct Chunk *next;
    size_t used, cap;
    char data[];
} Chunk;

typedef struct {
    Chunk *head;
} Arena;

static void *arena_alloc(Arena *a, size_t n) {
    n = (n + 15) & ~(size_t)15;
    if (!a->head || a->head->used + n > a->head->cap) {
        size_t cap = (size_t)1 << 22;
        if (cap < n) cap = n;
        Chunk *c = (Chunk *)malloc(sizeof(Chunk) + cap);
        if (!c) abort();
        c->next = a->head;
        c->used = 0;
        c->cap = cap;
        a->head = c;
    }
    void *p = a->head->data + a->head->used;
    a->head->used += n;
    memset(p, 0, n);
    return p;
}

static void arena_free(Arena *a) {
    Chunk *c = a->head;
    while (c) {
        Chunk *n = c->next;
        free(c);
        c = n;
    }
    a->head = NULL;
}

struct ora_matcher {
    Arena arena;
    Node *root;
    int family;
    int caseSensitive;
    int is_map;
    int32_t charBufferSize;
    int64_t nodes;
    uint8_t *wordChars; /* WholeWord only */
};

/* HashmapNode(boolean root[, int level]) — AhoCorasickSet.java:262-272 */
static Node *new_hashmap_node(ora_matcher *m, int root, int32_t level) {
    Node *n = (Node *)arena_alloc(&m->arena, sizeof(Node));
    n->kind = 0;
    n->children = (Node **)arena_alloc(&m->arena, sizeof(Node *));
    n->keys = (uint16_t *)arena_alloc(&m->arena, sizeof(uint16_t));
    n->cap = 1;
    n->modulusMask = 0;
    n->numEntries = 0;
    n->defaultTransition = root ? n : NULL;
    n->level = level;
    n->value = -1;
    n->failValue = -1;
    m->nodes++;
    return n;
}

/* FNV-1a over the two bytes of the char — AhoCorasickSet.java:406-410 */
static inline uint32_t hash_char(uint16_t c) {
    const uint32_t HASH_PRIME = 16777619u;
    return (((0x811c9dc5u ^ (uint32_t)(c >> 8)) * HASH_PRIME) ^ (uint32_t)(c & 0xff)) * HASH_PRIME;
}

/* HashmapNode.getTransition AhoCorasickSet.java:275-289; RangeNode.getTransition :451-460 */
static inline Node *get_transition(const Node *n, uint16_t key) {
    if (n->kind == 0) {
        int32_t defaultSlot = (int32_t)(hash_char(key) & (uint32_t)n->modulusMask);
        int32_t currentSlot = defaultSlot;
        do {
            if (n->keys[currentSlot] == key) {
                return n->children[currentSlot];
            } else if (n->children[currentSlot] == NULL) {
                return n->defaultTransition;
            } else {
                currentSlot = (currentSlot + 1) & n->modulusMask;
            }
        } while (currentSlot != defaultSlot);
        return n->defaultTransition;
    } else {
        int32_t idx = (uint16_t)(key - n->baseChar);
        if (idx < n->size) {
            return n->children[idx];
        }
        return n->defaultTransition;
    }
}

static inline int is_empty(const Node *n) { return n->kind == 0 ? n->numEntries == 0 : n->size == 0; }

/* HashmapNode.enlarge — AhoCorasickSet.java:350-376 */
static void enlarge(ora_matcher *m, Node *n) {
    int32_t ncap = n->cap * 2;
    uint16_t *biggerKeys = (uint16_t *)arena_alloc(&m->arena, sizeof(uint16_t) * (size_t)ncap);
    Node **biggerChildren = (Node **)arena_alloc(&m->arena, sizeof(Node *) * (size_t)ncap);
    int32_t biggerMask = ncap - 1;
    for (int32_t i = 0; i < n->cap; i++) {
        uint16_t key = n->keys[i];
        Node *node = n->children[i];
        if (node != NULL) {
            int32_t defaultSlot = (int32_t)(hash_char(key) & (uint32_t)biggerMask);
            int32_t currentSlot = defaultSlot;
            do {
                if (biggerChildren[currentSlot] == NULL) {
                    biggerKeys[currentSlot] = key;
                    biggerChildren[currentSlot] = node;
                    break;
                } else {
                    currentSlot = (currentSlot + 1) & biggerMask;
                }
            } while (currentSlot != defaultSlot);
        }
    }
    n->keys = biggerKeys;
    n->children = biggerChildren;
    n->modulusMask = biggerMask;
    n->cap = ncap;
}

/* HashmapNode.getOrAddChild — AhoCorasickSet.java:380-403 (level+1: LongestMatchSet.java:405) */
static Node *get_or_add_child(ora_matcher *m, Node *n, uint16_t key) {
    if (n->cap < 0x10000 &&
        ((n->numEntries >= n->cap) || (n->numEntries > 16 && ((float)n->numEntries >= (float)n->cap * 0.90f)))) {
        enlarge(m, n);
    }
    int32_t defaultSlot = (int32_t)(hash_char(key) & (uint32_t)n->modulusMask);
    int32_t currentSlot = defaultSlot;
    do {
        if (n->children[currentSlot] == NULL) {
            n->keys[currentSlot] = key;
            Node *newChild = new_hashmap_node(m, 0, n->level + 1);
            n->children[currentSlot] = newChild;
            ++n->numEntries;
            return newChild;
        } else if (n->keys[currentSlot] == key) {
            return n->children[currentSlot];
        } else {
            currentSlot = (currentSlot + 1) & n->modulusMask;
        }
    } while (currentSlot != defaultSlot);
    abort(); /* IllegalStateException */
}

/* RangeNodeThreshold.isOverThreshold with the default parameters
 * (exponent 1, linearFactor 1, maxValue 0.65, constantFactor 2) — threshold/RangeNodeThreshold.java:7-29 */
static int is_over_threshold(int32_t nodeSize, int32_t nodeLevel, int32_t keyIntervalSize) {
    if (keyIntervalSize <= 8) {
        return 1;
    }
    int32_t charArrayCost = (nodeSize / 4) + 3;
    return (double)(nodeSize + charArrayCost) > (double)keyIntervalSize * (0.65 - 1.0 / pow(2.0 + (double)nodeLevel, 1.0));
}

/* HashmapNode.optimizeNode + RangeNode ctor — AhoCorasickSet.java:323-346, :423-446.
 * The Java code allocates a new RangeNode and re-points the parent; nothing else
 * references the old node at that moment, so converting in place is equivalent.
 * force_if_default: AC/Longest/Shortest turn the root (defaultTransition != null) into a
 * RangeNode unconditionally; WholeWord has no default transition (WholeWordMatchSet.java:282-287). */
static void optimize_node(ora_matcher *m, Node *n, int32_t level) {
    if (n->kind != 0) return;
    uint16_t minKey = 0xffff;
    uint16_t maxKey = 0;
    int32_t size = n->numEntries;
    for (int32_t i = 0; i < n->cap; i++) {
        if (n->children[i] != NULL) {
            if (n->keys[i] > maxKey) maxKey = n->keys[i];
            if (n->keys[i] < minKey) minKey = n->keys[i];
        }
    }
    int32_t keyIntervalSize = (int32_t)maxKey - (int32_t)minKey + 1;
    if (n->defaultTransition != NULL || is_over_threshold(size, level, keyIntervalSize)) {
        int32_t rsize = (int32_t)maxKey - (int32_t)minKey + 1;
        Node **rchildren = NULL;
        if (rsize <= 0) {
            rsize = 0;
        } else {
            rchildren = (Node **)arena_alloc(&m->arena, sizeof(Node *) * (size_t)rsize);
            if (n->defaultTransition != NULL) {
                for (int32_t i = 0; i < rsize; i++) rchildren[i] = n;
            }
            for (int32_t i = 0; i < n->cap; i++) {
                if (n->children[i] != NULL) {
                    rchildren[n->keys[i] - minKey] = n->children[i];
                }
            }
        }
        n->kind = 1;
        n->baseChar = minKey;
        n->size = rsize;
        n->children = rchildren;
        n->keys = NULL;
    }
}

/* ShortestMatchSet.java:283-288 (HashmapNode.clear), :464-468 (RangeNode.clear) */
static void clear_node(ora_matcher *m, Node *n) {
    if (n->kind == 0) {
        n->children = (Node **)arena_alloc(&m->arena, sizeof(Node *));
        n->keys = (uint16_t *)arena_alloc(&m->arena, sizeof(uint16_t));
        n->cap = 1;
        n->modulusMask = 0;
        n->numEntries = 0;
    } else {
        n->children = NULL;
        n->size = 0;
    }
}

/* Queue.java: push / take (FIFO) / pop (LIFO). NULL is a legal element (level marker). */
typedef struct {
    Node **a;
    size_t first, last, cap;
} NQueue;

static void q_push(NQueue *q, Node *n) {
    if (q->last == q->cap) {
        if (q->first > 0 && q->first >= q->cap / 2) {
            memmove(q->a, q->a + q->first, (q->last - q->first) * sizeof(Node *));
            q->last -= q->first;
            q->first = 0;
        } else {
            q->cap = q->cap ? q->cap * 2 : 64;
            q->a = (Node **)realloc(q->a, q->cap * sizeof(Node *));
            if (!q->a) abort();
        }
    }
    q->a[q->last++] = n;
}
static int q_empty(const NQueue *q) { return q->first == q->last; }
static Node *q_take(NQueue *q) { return q_empty(q) ? NULL : q->a[q->first++]; }
static Node *q_pop(NQueue *q) { return q_empty(q) ? NULL : q->a[--q->last]; }

/* mapEntries order: HashmapNode by slot (AhoCorasickSet.java:297-303), RangeNode by char (:468-476).
 * Collects (key, child) pairs so a visitor may mutate the child (never the parent's key set). */
typedef void (*visit_fn)(ora_matcher *m, NQueue *q, Node *parent, uint16_t key, Node *value, int32_t level);

static void map_entries(ora_matcher *m, NQueue *q, Node *n, visit_fn visit, int32_t level) {
    if (n->kind == 0) {
        for (int32_t i = 0; i < n->cap; i++) {
            if (n->children[i] != NULL) {
                visit(m, q, n, n->keys[i], n->children[i], level);
            }
        }
    } else if (n->children != NULL) {
        for (int32_t i = 0; i < n->size; i++) {
            if (n->children[i] != NULL && n->children[i] != n) {
                visit(m, q, n, (uint16_t)(n->baseChar + i), n->children[i], level);
            }
        }
    }
}

/* Fail-transition search shared by AC / Longest / Shortest — AhoCorasickSet.java:64-91 */
static void compute_fail(Node *parent, uint16_t key, Node *value) {
    Node *parentFail = parent->failTransition;
    if (parentFail == NULL) {
        value->failTransition = parent;
    } else {
        do {
            Node *matchContinuation = get_transition(parentFail, key);
            if (matchContinuation != NULL) {
                value->failTransition = matchContinuation;
            } else {
                parentFail = parentFail->failTransition;
            }
        } while (value->failTransition == NULL);
    }
}

/* failTransAndOutputsVisitor of AhoCorasickSet.java:56-128 / AhoCorasickMap.java:70-145 and
 * LongestMatchSet.java:55-127 (identical apart from `level`). */
static void visit_ac(ora_matcher *m, NQueue *q, Node *parent, uint16_t key, Node *value, int32_t level) {
    optimize_node(m, value, level);
    int hadParentFail = parent->failTransition != NULL;
    compute_fail(parent, key, value);
    if (hadParentFail) {
        Node *fail = value->failTransition;
        while (fail != m->root && fail->matchLength == 0) {
            fail = fail->failTransition;
        }
        if (fail->matchLength > 0) {
            if (value->matchLength == 0) {
                value->matchLength = fail->matchLength;
                value->suffixMatch = fail->suffixMatch;
                value->value = fail->value; /* AhoCorasickMap.java:132 */
            } else {
                value->suffixMatch = fail;
            }
        }
    }
    if (!is_empty(value)) {
        q_push(q, value);
    }
}

/* failTransAndOutputsVisitor of ShortestMatchSet.java:59-116 / ShortestMatchMap.java:77-136 */
static void visit_shortest(ora_matcher *m, NQueue *q, Node *parent, uint16_t key, Node *value, int32_t level) {
    optimize_node(m, value, level);
    int hadParentFail = parent->failTransition != NULL;
    compute_fail(parent, key, value);
    if (hadParentFail) {
        if (value->matchLength == 0) {
            Node *fail = value->failTransition;
            while (fail != m->root && fail->matchLength == 0) {
                fail = fail->failTransition;
            }
            value->matchLength = fail->matchLength;
            value->value = fail->value; /* ShortestMatchMap.java:118 */
        }
        if (value->matchLength != 0) {
            clear_node(m, value);
            value->failTransition = m->root;
        }
    }
    if (!is_empty(value)) {
        q_push(q, value);
    }
}

/* optimizeNodesVisitor of WholeWordMatchSet.java:183-190 — note: it never enqueues the child,
 * so only the root's children are ever optimised. */
static void visit_wholeword(ora_matcher *m, NQueue *q, Node *parent, uint16_t key, Node *value, int32_t level) {
    (void)q;
    (void)parent;
    (void)key;
    /* WholeWord HashmapNode.optimizeNode has no defaultTransition clause (WholeWordMatchSet.java:282-287);
     * defaultTransition is always NULL for this family so optimize_node behaves identically. */
    optimize_node(m, value, level);
}

/* optimizeNodesAndFailTransitions of WholeWordLongestMatchSet.java:224-247 / WholeWordLongestMatchMap.java:362-386:
 * the fail match of a node is the deepest keyword on its path that is followed (on the path) by a non-word char. */
static void visit_wwlongest(ora_matcher *m, NQueue *q, Node *parent, uint16_t key, Node *value, int32_t level) {
    optimize_node(m, value, level);
    if (parent->matchLength != 0 && !m->wordChars[key]) {
        value->failMatchLength = parent->matchLength;
        value->failMatchOffset = 1;
        value->failValue = parent->value;
    } else {
        value->failMatchLength = parent->failMatchLength;
        value->failMatchOffset = parent->failMatchOffset + 1;
        value->failValue = parent->failValue;
    }
    if (!is_empty(value)) {
        q_push(q, value);
    }
}

/* enqueueNodesVisitor — AhoCorasickSet.java:147-155 */
static void visit_enqueue(ora_matcher *m, NQueue *q, Node *parent, uint16_t key, Node *value, int32_t level) {
    (void)m;
    (void)parent;
    (void)key;
    (void)level;
    if (!is_empty(value)) {
        q_push(q, value);
    }
}

/* The breadth-first driver — AhoCorasickSet.java:49-53,130-140 */
static void bfs(ora_matcher *m, NQueue *q, visit_fn visit) {
    optimize_node(m, m->root, 0);
    q_push(q, m->root);
    q_push(q, NULL);
    int32_t level = 1;
    while (!q_empty(q)) {
        Node *n = q_take(q);
        if (n == NULL) {
            if (!q_empty(q)) {
                q_push(q, NULL);
                level++;
            }
        } else {
            map_entries(m, q, n, visit, level);
        }
    }
}

/* RangeNode gap fill, depth first — AhoCorasickSet.java:156-190 (AC family only).
 * Restated literally, including the fact that after a null marker it pops the *next*
 * stack entry, so only part of the trie is visited; the filled transitions equal what the
 * match-time fail walk computes, so results do not depend on it. */
static void gap_fill(ora_matcher *m, NQueue *q) {
    map_entries(m, q, m->root, visit_enqueue, 0);
    while (!q_empty(q)) {
        Node *node = q_pop(q);
        if (node == NULL) {
            node = q_pop(q);
            if (node != NULL && node->kind == 1) {
                Node *rangeNode = node;
                for (int32_t i = 0; i < rangeNode->size; i++) {
                    if (rangeNode->children[i] == NULL) {
                        uint16_t charOfMissingTransition = (uint16_t)(rangeNode->baseChar + i);
                        Node *n = rangeNode->failTransition;
                        while (n != NULL) {
                            Node *nextNode = get_transition(n, charOfMissingTransition);
                            if (nextNode == NULL) {
                                n = n->failTransition;
                            } else {
                                rangeNode->children[i] = nextNode;
                                break;
                            }
                        }
                    }
                }
            }
        } else {
            q_push(q, NULL);
            map_entries(m, q, node, visit_enqueue, 0);
        }
    }
}

/* WordCharacters.trim — WordCharacters.java:41-62. Returns [*ws, *we). */
static void trim_keyword(const uint16_t *kw, int32_t len, const uint8_t *wordChars, int32_t *ws, int32_t *we) {
    int32_t wordStart = 0;
    int32_t wordEnd = len;
    for (int32_t i = 0; i < len; i++) {
        if (wordChars[kw[i]]) {
            wordStart = i;
            break;
        }
    }
    for (int32_t i = len - 1; i >= 0; i--) {
        if (wordChars[kw[i]]) {
            wordEnd = i + 1;
            break;
        }
    }
    *ws = wordStart;
    *we = wordEnd;
}

static void set_err_nonword(char *err, int errlen, const uint16_t *kw, int32_t len) {
    if (!err || errlen <= 0) return;
    int p = 0;
    for (int32_t i = 0; i < len && p < errlen - 40; i++) {
        uint16_t c = kw[i];
        if (c == 0) { /* modified UTF-8, so the C string is not cut short */
            err[p++] = (char)0xC0;
            err[p++] = (char)0x80;
        } else if (c < 0x80) {
            err[p++] = (char)c;
        } else if (c < 0x800) {
            err[p++] = (char)(0xC0 | (c >> 6));
            err[p++] = (char)(0x80 | (c & 0x3F));
        } else {
            err[p++] = (char)(0xE0 | (c >> 12));
            err[p++] = (char)(0x80 | ((c >> 6) & 0x3F));
            err[p++] = (char)(0x80 | (c & 0x3F));
        }
    }
    snprintf(err + p, (size_t)(errlen - p), " contains non-word characters.");
}

ora_matcher *ora_create(int family, const uint16_t *chars, const int64_t *offsets,
                        const uint8_t *is_null, int64_t n_keywords, int64_t n_values,
                        int case_sensitive, const uint8_t *word_chars,
                        char *err, int errlen) {
    ensure_tables();
    if (err && errlen > 0) err[0] = 0;
    ora_matcher *m = (ora_matcher *)calloc(1, sizeof(ora_matcher));
    m->family = family;
    m->caseSensitive = case_sensitive != 0;
    m->is_map = n_values >= 0;
    int cs = m->caseSensitive;
    /* Maps zip keywords with values and stop at the shorter (AhoCorasickMap.java:32). */
    int64_t n = n_keywords;
    if (m->is_map && n_values < n) n = n_values;

    if (family == ORA_WHOLEWORD || family == ORA_WHOLEWORDLONGEST) {
        m->wordChars = (uint8_t *)malloc(65536);
        if (word_chars) {
            memcpy(m->wordChars, word_chars, 65536);
        } else {
            ora_word_chars(0, NULL, NULL, 0, m->wordChars);
        }
        /* root = new HashmapNode()  — no default transition (WholeWordMatchSet.java:142) */
        m->root = new_hashmap_node(m, 0, 0);
    } else {
        m->root = new_hashmap_node(m, 1, 0);
    }

    int32_t longestKeyword = 0;
    for (int64_t k = 0; k < n; k++) {
        if (is_null && is_null[k]) continue;
        const uint16_t *kw = chars + offsets[k];
        int32_t len = (int32_t)(offsets[k + 1] - offsets[k]);
        if (family == ORA_WHOLEWORDLONGEST) {
            /* WholeWordLongestMatchSet.java:190-206 / WholeWordLongestMatchMap.java:322-344: trim, no validation */
            int32_t ws, we;
            trim_keyword(kw, len, m->wordChars, &ws, &we);
            kw += ws;
            len = we - ws;
            if (len > 0) {
                if (len > longestKeyword) longestKeyword = len;
                Node *currentNode = m->root;
                for (int32_t idx = 0; idx < len; idx++) {
                    currentNode = get_or_add_child(m, currentNode, cs ? kw[idx] : g_lower[kw[idx]]);
                }
                currentNode->matchLength = len;
                currentNode->value = m->is_map ? (int32_t)k : -1;
            }
        } else if (family == ORA_WHOLEWORD) {
            /* WholeWordMatchSet.java:145-166 / WholeWordMatchMap.java:262-285 */
            int32_t ws, we;
            trim_keyword(kw, len, m->wordChars, &ws, &we);
            kw += ws;
            len = we - ws;
            for (int32_t i = 0; i < len; i++) {
                if (!m->wordChars[kw[i]]) {
                    set_err_nonword(err, errlen, kw, len);
                    ora_destroy(m);
                    return NULL;
                }
            }
            if (len > longestKeyword) longestKeyword = len;
            if (len > 0) {
                Node *currentNode = m->root;
                for (int32_t idx = 0; idx < len; idx++) {
                    currentNode = get_or_add_child(m, currentNode, cs ? kw[idx] : g_lower[kw[idx]]);
                }
                currentNode->matchLength = len;
                currentNode->value = m->is_map ? (int32_t)k : -1;
            }
        } else if (len > 0) {
            if (len > longestKeyword) longestKeyword = len;
            Node *currentNode = m->root;
            int pruned = 0;
            for (int32_t idx = 0; idx < len; idx++) {
                currentNode = get_or_add_child(m, currentNode, cs ? kw[idx] : g_lower[kw[idx]]);
                /* ShortestMatchSet.java:32-36: a keyword with an earlier-inserted prefix (or equal) keyword
                 * is dropped — `continue OUTER` before matchLength/value are assigned. */
                if (family == ORA_SHORTEST && currentNode->matchLength != 0) {
                    pruned = 1;
                    break;
                }
            }
            if (!pruned) {
                currentNode->matchLength = len;
                currentNode->value = m->is_map ? (int32_t)k : -1;
            }
        }
    }
    m->charBufferSize = longestKeyword > 2048 ? longestKeyword * 2 : 4096;

    NQueue q = {0};
    switch (family) {
    case ORA_AHOCORASICK:
        bfs(m, &q, visit_ac);
        gap_fill(m, &q);
        break;
    case ORA_LONGEST:
        bfs(m, &q, visit_ac);
        break;
    case ORA_SHORTEST:
        bfs(m, &q, visit_shortest);
        break;
    case ORA_WHOLEWORD:
        bfs(m, &q, visit_wholeword);
        break;
    case ORA_WHOLEWORDLONGEST:
        bfs(m, &q, visit_wwlongest);
        break;
    default:
        free(q.a);
        ora_destroy(m);
        return NULL;
    }
    free(q.a);
    return m;
}

void ora_destroy(ora_matcher *m) {
    if (!m) return;
    arena_free(&m->arena);
    free(m->wordChars);
    free(m);
}

int32_t ora_char_buffer_size(const ora_matcher *m) { return m->charBufferSize; }
int64_t ora_node_count(const ora_matcher *m) { return m->nodes; }

/* ------------------------------------------------------------------ listener plumbing */

typedef struct {
    ora_listener cb;
    void *ctx;
    int64_t calls;
} Sink;

static inline int emit(Sink *s, int32_t start, int32_t end, int32_t value) {
    s->calls++;
    return s->cb(s->ctx, start, end, value);
}

/* ------------------------------------------------------------------ SetMatchQueue / MapMatchQueue */

struct ora_queue {
    int32_t emptySlotIdx;
    int32_t length; /* endIndexes.length */
    int32_t *endIndexes;
    int32_t *startIndexes;
    int32_t *values;
};

static void queue_init(ora_queue *q) {
    q->emptySlotIdx = 0;
    q->length = 2;
    q->endIndexes = (int32_t *)malloc(2 * sizeof(int32_t));
    q->startIndexes = (int32_t *)malloc(2 * sizeof(int32_t));
    q->values = (int32_t *)malloc(2 * sizeof(int32_t));
}

static void queue_release(ora_queue *q) {
    free(q->endIndexes);
    free(q->startIndexes);
    free(q->values);
}

/* SetMatchQueue.matchAndClear — SetMatchQueue.java:19-42 (MapMatchQueue.java:21-72).
 * readable != 0: values only. Returns 0 when the listener answered false. */
static int queue_match_and_clear(ora_queue *q, Sink *s, int32_t purgeToIndex, int readable) {
    if (q->emptySlotIdx != 0) {
        int32_t i = 0;
        while (i < q->emptySlotIdx) {
            if (q->endIndexes[i] <= purgeToIndex) {
                int ok = readable ? emit(s, -1, -1, q->values[i])
                                  : emit(s, q->startIndexes[i], q->endIndexes[i], q->values[i]);
                if (!ok) {
                    return 0;
                }
            } else {
                break;
            }
            i++;
        }
        if (i > 0) {
            q->emptySlotIdx = q->emptySlotIdx - i;
            memmove(q->endIndexes, q->endIndexes + i, (size_t)q->emptySlotIdx * sizeof(int32_t));
            memmove(q->startIndexes, q->startIndexes + i, (size_t)q->emptySlotIdx * sizeof(int32_t));
            memmove(q->values, q->values + i, (size_t)q->emptySlotIdx * sizeof(int32_t));
        }
    }
    return 1;
}

/* SetMatchQueue.push — SetMatchQueue.java:45-95 (MapMatchQueue.java:75-132) */
static int queue_push(ora_queue *q, int32_t length, int32_t idx, int32_t value) {
    if (q->emptySlotIdx + 1 == q->length) {
        int32_t newCapacity = q->length * 2;
        q->endIndexes = (int32_t *)realloc(q->endIndexes, (size_t)newCapacity * sizeof(int32_t));
        q->startIndexes = (int32_t *)realloc(q->startIndexes, (size_t)newCapacity * sizeof(int32_t));
        q->values = (int32_t *)realloc(q->values, (size_t)newCapacity * sizeof(int32_t));
        q->length = newCapacity;
    }
    if (q->emptySlotIdx != 0) {
        int32_t idxToFind = idx - length;
        for (int32_t currSlot = q->emptySlotIdx - 1; currSlot >= 0; currSlot--) {
            int32_t currStartIdx = q->startIndexes[currSlot];
            if (idxToFind >= currStartIdx) {
                if (idxToFind >= q->endIndexes[currSlot]) {
                    q->startIndexes[currSlot + 1] = idxToFind;
                    q->endIndexes[currSlot + 1] = idx;
                    q->values[currSlot + 1] = value;
                    q->emptySlotIdx = currSlot + 2;
                    return 1;
                } else if (idxToFind == currStartIdx && q->endIndexes[currSlot] < idx) {
                    q->startIndexes[currSlot] = idxToFind;
                    q->endIndexes[currSlot] = idx;
                    q->values[currSlot] = value;
                    q->emptySlotIdx = currSlot + 1;
                    return 1;
                } else {
                    return 0;
                }
            }
        }
        q->startIndexes[0] = idxToFind;
        q->endIndexes[0] = idx;
        q->values[0] = value;
        q->emptySlotIdx = 1;
        return 1;
    } else {
        q->startIndexes[q->emptySlotIdx] = idx - length;
        q->endIndexes[q->emptySlotIdx] = idx;
        q->values[q->emptySlotIdx] = value;
        q->emptySlotIdx++;
        return 1;
    }
}

ora_queue *ora_queue_new(void) {
    ora_queue *q = (ora_queue *)malloc(sizeof(ora_queue));
    queue_init(q);
    return q;
}
void ora_queue_free(ora_queue *q) {
    if (!q) return;
    queue_release(q);
    free(q);
}
int ora_queue_push(ora_queue *q, int32_t length, int32_t idx) { return queue_push(q, length, idx, -1); }

typedef struct {
    ora_match *out;
    int64_t cap;
    int64_t n;
    int64_t stop_after;
} Collector;

static int collect_cb(void *ctx, int32_t start, int32_t end, int32_t value) {
    Collector *c = (Collector *)ctx;
    if (c->n < c->cap) {
        c->out[c->n].start = start;
        c->out[c->n].end = end;
        c->out[c->n].value = value;
    }
    c->n++;
    if (c->stop_after > 0 && c->n >= c->stop_after) return 0;
    return 1;
}

int64_t ora_queue_match_and_clear(ora_queue *q, int32_t purge_to, ora_match *out, int64_t cap) {
    Collector c = {out, cap, 0, 0};
    Sink s = {collect_cb, &c, 0};
    queue_match_and_clear(q, &s, purge_to, 0);
    return c.n;
}

/* ------------------------------------------------------------------ CharBuffer + Readable emulation */

typedef struct {
    uint16_t *a;
    int32_t cap, pos, lim;
} CharBuf;

typedef struct {
    const uint16_t *data;
    int64_t n, at;
    const int32_t *schedule;
    int64_t n_schedule, k;
} Reader;

/* Readable.read(CharBuffer): -1 at end of input, else number of chars put. */
static int32_t reader_read(Reader *r, CharBuf *b) {
    if (r->at >= r->n) return -1;
    int64_t want = b->lim - b->pos;
    if (r->n - r->at < want) want = r->n - r->at;
    if (r->schedule && r->k < r->n_schedule) {
        if (r->schedule[r->k] < want) want = r->schedule[r->k] < 0 ? 0 : r->schedule[r->k];
        r->k++;
    }
    memcpy(b->a + b->pos, r->data + r->at, (size_t)want * sizeof(uint16_t));
    b->pos += (int32_t)want;
    r->at += want;
    return (int32_t)want;
}
static inline void buf_flip(CharBuf *b) {
    b->lim = b->pos;
    b->pos = 0;
}
static inline void buf_clear(CharBuf *b) {
    b->pos = 0;
    b->lim = b->cap;
}
static inline int buf_has_remaining(const CharBuf *b) { return b->pos < b->lim; }
static inline uint16_t buf_get(CharBuf *b) { return b->a[b->pos++]; }

/* ------------------------------------------------------------------ AhoCorasick */

/* TrieNode.output(String, listener, idx) — AhoCorasickSet.java:522-535 / AhoCorasickMap.java:627-640;
 * readable: output(ReadableMatchListener) AhoCorasickMap.java:611-624 */
static inline int ac_output(const Node *n, Sink *s, int32_t idx, int readable) {
    int ret = 1;
    if (n->matchLength > 0) {
        ret = readable ? emit(s, -1, -1, n->value) : emit(s, idx - n->matchLength, idx, n->value);
        const Node *suffixMatch = n->suffixMatch;
        while (suffixMatch != NULL && ret) {
            ret = readable ? emit(s, -1, -1, suffixMatch->value)
                           : emit(s, idx - suffixMatch->matchLength, idx, suffixMatch->value);
            suffixMatch = suffixMatch->suffixMatch;
        }
    }
    return ret;
}

/* AhoCorasickSet.match — AhoCorasickSet.java:193-252 (cs const-folds into the two Java copies) */
static inline __attribute__((always_inline)) void ac_match_string(const ora_matcher *m, const uint16_t *haystack,
                                                                 int32_t len, Sink *s, const int cs) {
    const Node *currentNode = m->root;
    int32_t idx = 0;
    while (idx < len) {
        const uint16_t c = cs ? haystack[idx] : g_lower[haystack[idx]];
        const Node *nextNode = get_transition(currentNode, c);
        while (nextNode == NULL) {
            currentNode = currentNode->failTransition;
            nextNode = get_transition(currentNode, c);
        }
        currentNode = nextNode;
        if (!ac_output(currentNode, s, ++idx, 0)) {
            break;
        }
    }
}

/* AhoCorasickMap.match(Readable) — AhoCorasickMap.java:208-275 */
static void ac_match_readable(const ora_matcher *m, Reader *haystack, Sink *s, const int cs) {
    const Node *currentNode = m->root;
    CharBuf buf = {(uint16_t *)malloc((size_t)m->charBufferSize * 2), m->charBufferSize, 0, m->charBufferSize};
    while (reader_read(haystack, &buf) != -1) {
        buf_flip(&buf);
        while (buf_has_remaining(&buf)) {
            uint16_t c = buf_get(&buf);
            if (!cs) c = g_lower[c];
            const Node *nextNode = get_transition(currentNode, c);
            while (nextNode == NULL) {
                currentNode = currentNode->failTransition;
                nextNode = get_transition(currentNode, c);
            }
            currentNode = nextNode;
            if (!ac_output(currentNode, s, 0, 1)) {
                free(buf.a);
                return;
            }
        }
        buf_clear(&buf);
    }
    free(buf.a);
}

/* ------------------------------------------------------------------ LongestMatch */

/* TrieNode.output(queue, idx) — LongestMatchSet.java:535-551 / LongestMatchMap.java:635-653 */
static inline void longest_output(const Node *n, ora_queue *queue, int32_t idx) {
    int matchAccepted = 0;
    if (n->matchLength != 0) {
        matchAccepted = queue_push(queue, n->matchLength, idx, n->value);
        const Node *suffixMatch = n->suffixMatch;
        while (suffixMatch != NULL && !matchAccepted) {
            matchAccepted = queue_push(queue, suffixMatch->matchLength, idx, suffixMatch->value);
            suffixMatch = suffixMatch->suffixMatch;
        }
    }
}

/* LongestMatchSet.match — LongestMatchSet.java:192-265 */
static inline __attribute__((always_inline)) void longest_match_string(const ora_matcher *m, const uint16_t *haystack,
                                                                      int32_t len, Sink *s, const int cs) {
    const Node *currentNode = m->root;
    ora_queue queue;
    queue_init(&queue);
    int32_t idx = 0;
    while (idx < len) {
        const uint16_t c = cs ? haystack[idx] : g_lower[haystack[idx]];
        const Node *nextNode = get_transition(currentNode, c);
        int failTransition = 0;
        while (nextNode == NULL) {
            failTransition = 1;
            currentNode = currentNode->failTransition;
            nextNode = get_transition(currentNode, c);
        }
        currentNode = nextNode;
        longest_output(currentNode, &queue, ++idx);
        if (failTransition && !queue_match_and_clear(&queue, s, idx - currentNode->level, 0)) {
            queue_release(&queue);
            return;
        }
    }
    queue_match_and_clear(&queue, s, INT32_MAX, 0);
    queue_release(&queue);
}

/* LongestMatchMap.match(Readable) — LongestMatchMap.java:203-286 */
static void longest_match_readable(const ora_matcher *m, Reader *haystack, Sink *s, const int cs) {
    const Node *currentNode = m->root;
    ora_queue queue;
    queue_init(&queue);
    CharBuf buf = {(uint16_t *)malloc((size_t)m->charBufferSize * 2), m->charBufferSize, 0, m->charBufferSize};
    int32_t idx = 0;
    while (reader_read(haystack, &buf) != -1) {
        buf_flip(&buf);
        while (buf_has_remaining(&buf)) {
            uint16_t c = buf_get(&buf);
            if (!cs) c = g_lower[c];
            const Node *nextNode = get_transition(currentNode, c);
            int failTransition = 0;
            while (nextNode == NULL) {
                failTransition = 1;
                currentNode = currentNode->failTransition;
                nextNode = get_transition(currentNode, c);
            }
            currentNode = nextNode;
            longest_output(currentNode, &queue, ++idx);
            if (failTransition && !queue_match_and_clear(&queue, s, idx - currentNode->level, 1)) {
                queue_release(&queue);
                free(buf.a);
                return;
            }
        }
        buf_clear(&buf);
    }
    queue_match_and_clear(&queue, s, INT32_MAX, 1);
    queue_release(&queue);
    free(buf.a);
}

/* ------------------------------------------------------------------ ShortestMatch */

/* ShortestMatchSet.match — ShortestMatchSet.java:182-260 / ShortestMatchMap.java:293-373.
 * Note the `break` on a false return falls into the post-loop emit (quirk Q1). */
static inline __attribute__((always_inline)) void shortest_match_string(const ora_matcher *m, const uint16_t *haystack,
                                                                       int32_t len, Sink *s, const int cs) {
    const Node *root = m->root;
    const Node *currentNode = root;
    int32_t currentNodeMatchLength = currentNode->matchLength;
    int32_t currentNodeMatchValue = currentNode->value;
    int32_t idx = 0;
    while (idx < len) {
        const uint16_t c = cs ? haystack[idx] : g_lower[haystack[idx]];
        if (currentNodeMatchLength != 0) {
            if (!emit(s, idx - currentNodeMatchLength, idx, currentNodeMatchValue)) {
                break;
            }
            currentNode = get_transition(root, c);
        } else {
            const Node *nextNode = get_transition(currentNode, c);
            while (nextNode == NULL) {
                currentNode = currentNode->failTransition;
                nextNode = get_transition(currentNode, c);
            }
            currentNode = nextNode;
        }
        currentNodeMatchLength = currentNode->matchLength;
        currentNodeMatchValue = currentNode->value;
        ++idx;
    }
    if (currentNodeMatchLength != 0) {
        emit(s, idx - currentNodeMatchLength, idx, currentNodeMatchValue);
    }
}

/* ShortestMatchMap.match(Readable) — ShortestMatchMap.java:199-291.
 * The pending match is emitted at the end of *every* buffer fill and not cleared (quirk Q4). */
static void shortest_match_readable(const ora_matcher *m, Reader *haystack, Sink *s, const int cs) {
    const Node *root = m->root;
    CharBuf buf = {(uint16_t *)malloc((size_t)m->charBufferSize * 2), m->charBufferSize, 0, m->charBufferSize};
    const Node *currentNode = root;
    int32_t currentNodeMatchLength = currentNode->matchLength;
    int32_t currentNodeMatchValue = currentNode->value;
    while (reader_read(haystack, &buf) != -1) {
        buf_flip(&buf);
        while (buf_has_remaining(&buf)) {
            uint16_t c = buf_get(&buf);
            if (!cs) c = g_lower[c];
            if (currentNodeMatchLength != 0) {
                if (!emit(s, -1, -1, currentNodeMatchValue)) {
                    free(buf.a);
                    return;
                }
                currentNode = get_transition(root, c);
            } else {
                const Node *nextNode = get_transition(currentNode, c);
                while (nextNode == NULL) {
                    currentNode = currentNode->failTransition;
                    nextNode = get_transition(currentNode, c);
                }
                currentNode = nextNode;
            }
            currentNodeMatchLength = currentNode->matchLength;
            currentNodeMatchValue = currentNode->value;
        }
        buf_clear(&buf);
        if (currentNodeMatchLength != 0) {
            if (!emit(s, -1, -1, currentNodeMatchValue)) {
                free(buf.a);
                return;
            }
        }
    }
    free(buf.a);
}

/* ------------------------------------------------------------------ WholeWordMatch */

/* WholeWordMatchSet.match — WholeWordMatchSet.java:47-132 / WholeWordMatchMap.java:155-240.
 * First word-char test uses the (possibly lower-cased) c; both scroll loops use the raw
 * haystack char (quirk Q7). */
static inline __attribute__((always_inline)) void wholeword_match_string(const ora_matcher *m, const uint16_t *haystack,
                                                                        int32_t len, Sink *s, const int cs) {
    const uint8_t *wordChars = m->wordChars;
    const Node *root = m->root;
    const Node *currentNode = root;
    int32_t idx = 0;
    while (idx < len) {
        uint16_t c = cs ? haystack[idx] : g_lower[haystack[idx]];
        const Node *nextNode = get_transition(currentNode, c);
        if (nextNode == NULL) {
            if (!wordChars[c]) {
                if (currentNode->matchLength != 0) {
                    if (!emit(s, idx - currentNode->matchLength, idx, currentNode->value)) {
                        return;
                    }
                }
            } else {
                while (++idx < len && wordChars[haystack[idx]]) {
                    ;
                }
            }
            while (++idx < len && !wordChars[haystack[idx]]) {
                ;
            }
            currentNode = root;
        } else {
            ++idx;
            currentNode = nextNode;
        }
    }
    if (currentNode->matchLength != 0) {
        emit(s, idx - currentNode->matchLength, idx, currentNode->value);
    }
}

/* WholeWordMatchMap.scroll — WholeWordMatchMap.java:325-339. Returns 1 at end of input. */
static int ww_scroll(const ora_matcher *m, Reader *haystack, CharBuf *buf, int wordChars, int caseSensitive) {
    do {
        while (buf_has_remaining(buf)) {
            uint16_t ch = buf_get(buf);
            if (!caseSensitive) ch = g_lower[ch];
            if ((m->wordChars[ch] != 0) != (wordChars != 0)) {
                buf->pos = buf->pos - 1;
                return 0;
            }
        }
        buf_clear(buf);
        if (reader_read(haystack, buf) == -1) {
            return 1;
        }
        buf_flip(buf);
    } while (1);
}

/* WholeWordMatchMap.match(Readable) — WholeWordMatchMap.java:55-153 */
static void wholeword_match_readable(const ora_matcher *m, Reader *haystack, Sink *s, const int cs) {
    const uint8_t *wordChars = m->wordChars;
    const Node *root = m->root;
    const Node *currentNode = root;
    CharBuf buf = {(uint16_t *)malloc((size_t)m->charBufferSize * 2), m->charBufferSize, 0, m->charBufferSize};
    while (reader_read(haystack, &buf) != -1) {
        buf_flip(&buf);
        while (buf_has_remaining(&buf)) {
            uint16_t c = buf_get(&buf);
            if (!cs) c = g_lower[c];
            const Node *nextNode = get_transition(currentNode, c);
            if (nextNode == NULL) {
                if (!wordChars[c]) {
                    if (currentNode->matchLength != 0) {
                        if (!emit(s, -1, -1, currentNode->value)) {
                            free(buf.a);
                            return;
                        }
                    }
                } else {
                    if (ww_scroll(m, haystack, &buf, 1, cs)) {
                        currentNode = root;
                        goto main_loop_end;
                    }
                }
                currentNode = root;
                if (ww_scroll(m, haystack, &buf, 0, cs)) {
                    goto main_loop_end;
                }
            } else {
                currentNode = nextNode;
            }
        }
        buf_clear(&buf);
    }
main_loop_end:
    if (currentNode->matchLength != 0) {
        emit(s, -1, -1, currentNode->value);
    }
    free(buf.a);
}

/* ------------------------------------------------------------------ WholeWordLongest */

/* WholeWordLongestMatchSet.match — WholeWordLongestMatchSet.java:47-182 / WholeWordLongestMatchMap.java:183-310 */
static inline __attribute__((always_inline)) void wwlongest_match_string(const ora_matcher *m, const uint16_t *haystack,
                                                                        int32_t len, Sink *s, const int cs) {
    const uint8_t *wordChars = m->wordChars;
    const Node *root = m->root;
    const Node *currentNode = root;
    int32_t idx = 0;
    while (idx < len) {
        uint16_t c = cs ? haystack[idx] : g_lower[haystack[idx]];
        const Node *nextNode = get_transition(currentNode, c);
        if (nextNode == NULL) {
            if (!wordChars[c]) {
                if (currentNode->matchLength != 0) {
                    if (!emit(s, idx - currentNode->matchLength, idx, currentNode->value)) {
                        return;
                    }
                } else if (currentNode->failMatchLength != 0) {
                    int32_t failMatchEnd = idx - currentNode->failMatchOffset;
                    if (!emit(s, failMatchEnd - currentNode->failMatchLength, failMatchEnd, currentNode->failValue)) {
                        return;
                    }
                }
            } else {
                if (currentNode->failMatchLength != 0) {
                    int32_t failMatchEnd = idx - currentNode->failMatchOffset;
                    if (!emit(s, failMatchEnd - currentNode->failMatchLength, failMatchEnd, currentNode->failValue)) {
                        return;
                    }
                }
                while (++idx < len && wordChars[haystack[idx]]) {
                    ;
                }
            }
            while (++idx < len && !wordChars[haystack[idx]]) {
                ;
            }
            currentNode = root;
        } else {
            ++idx;
            currentNode = nextNode;
        }
    }
    if (currentNode->matchLength != 0) {
        if (!emit(s, idx - currentNode->matchLength, idx, currentNode->value)) {
            return;
        }
    } else if (currentNode->failMatchLength != 0) {
        int32_t failMatchEnd = idx - currentNode->failMatchOffset;
        if (!emit(s, failMatchEnd - currentNode->failMatchLength, failMatchEnd, currentNode->failValue)) {
            return;
        }
    }
}

/* WholeWordLongestMatchMap.match(Readable) — WholeWordLongestMatchMap.java:54-181 (scroll: :401-414, same as the
 * WholeWordMatchMap one) */
static void wwlongest_match_readable(const ora_matcher *m, Reader *haystack, Sink *s, const int cs) {
    const uint8_t *wordChars = m->wordChars;
    const Node *root = m->root;
    const Node *currentNode = root;
    CharBuf buf = {(uint16_t *)malloc((size_t)m->charBufferSize * 2), m->charBufferSize, 0, m->charBufferSize};
    while (reader_read(haystack, &buf) != -1) {
        buf_flip(&buf);
        while (buf_has_remaining(&buf)) {
            uint16_t c = buf_get(&buf);
            if (!cs) c = g_lower[c];
            const Node *nextNode = get_transition(currentNode, c);
            if (nextNode == NULL) {
                if (!wordChars[c]) {
                    if (currentNode->matchLength != 0) {
                        if (!emit(s, -1, -1, currentNode->value)) {
                            free(buf.a);
                            return;
                        }
                    } else if (currentNode->failMatchLength != 0) {
                        if (!emit(s, -1, -1, currentNode->failValue)) {
                            free(buf.a);
                            return;
                        }
                    }
                } else {
                    if (currentNode->failMatchLength != 0) {
                        if (!emit(s, -1, -1, currentNode->failValue)) {
                            free(buf.a);
                            return;
                        }
                    }
                    if (ww_scroll(m, haystack, &buf, 1, cs)) {
                        currentNode = root;
                        goto main_loop_end;
                    }
                }
                currentNode = root;
                if (ww_scroll(m, haystack, &buf, 0, cs)) {
                    goto main_loop_end;
                }
            } else {
                currentNode = nextNode;
            }
        }
        buf_clear(&buf);
    }
main_loop_end:
    if (currentNode->matchLength != 0) {
        emit(s, -1, -1, currentNode->value);
    } else if (currentNode->failMatchLength != 0) {
        emit(s, -1, -1, currentNode->failValue);
    }
    free(buf.a);
}

/* ------------------------------------------------------------------ public entry points */

int64_t ora_match_string(const ora_matcher *m, const uint16_t *hay, int32_t n, ora_listener cb, void *ctx) {
    Sink s = {cb, ctx, 0};
    switch (m->family) {
    case ORA_AHOCORASICK:
        if (m->caseSensitive) ac_match_string(m, hay, n, &s, 1);
        else ac_match_string(m, hay, n, &s, 0);
        break;
    case ORA_LONGEST:
        if (m->caseSensitive) longest_match_string(m, hay, n, &s, 1);
        else longest_match_string(m, hay, n, &s, 0);
        break;
    case ORA_SHORTEST:
        if (m->caseSensitive) shortest_match_string(m, hay, n, &s, 1);
        else shortest_match_string(m, hay, n, &s, 0);
        break;
    case ORA_WHOLEWORD:
        if (m->caseSensitive) wholeword_match_string(m, hay, n, &s, 1);
        else wholeword_match_string(m, hay, n, &s, 0);
        break;
    case ORA_WHOLEWORDLONGEST:
        if (m->caseSensitive) wwlongest_match_string(m, hay, n, &s, 1);
        else wwlongest_match_string(m, hay, n, &s, 0);
        break;
    }
    return s.calls;
}

int64_t ora_match_readable(const ora_matcher *m, const uint16_t *hay, int64_t n,
                           const int32_t *schedule, int64_t n_schedule, ora_listener cb, void *ctx) {
    Sink s = {cb, ctx, 0};
    Reader r = {hay, n, 0, schedule, n_schedule, 0};
    switch (m->family) {
    case ORA_AHOCORASICK:
        ac_match_readable(m, &r, &s, m->caseSensitive);
        break;
    case ORA_LONGEST:
        longest_match_readable(m, &r, &s, m->caseSensitive);
        break;
    case ORA_SHORTEST:
        shortest_match_readable(m, &r, &s, m->caseSensitive);
        break;
    case ORA_WHOLEWORD:
        wholeword_match_readable(m, &r, &s, m->caseSensitive);
        break;
    case ORA_WHOLEWORDLONGEST:
        wwlongest_match_readable(m, &r, &s, m->caseSensitive);
        break;
    }
    return s.calls;
}

int64_t ora_match_collect(const ora_matcher *m, const uint16_t *hay, int64_t n, int readable,
                          int64_t stop_after, ora_match *out, int64_t cap) {
    Collector c = {out, cap, 0, stop_after};
    if (readable) {
        ora_match_readable(m, hay, n, NULL, 0, collect_cb, &c);
    } else {
        ora_match_string(m, hay, (int32_t)n, collect_cb, &c);
    }
    return c.n;
}

static int count_cb(void *ctx, int32_t start, int32_t end, int32_t value) {
    (void)start;
    (void)end;
    (void)value;
    ++*(volatile int64_t *)ctx;
    return 1;
}

int64_t ora_match_count(const ora_matcher *m, const uint16_t *hay, int32_t n) {
    int64_t count = 0;
    ora_match_string(m, hay, n, count_cb, &count);
    return count;
}
