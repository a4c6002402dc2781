"""Query-side NLP helpers: stop words (15 languages), synonym expansion, did-you-mean, natural-language filter
parsing, related-search tracking.  Behaviour per reference infomesh/search/nlp.py:741-1005.
"""
from __future__ import annotations

import re
import time
from collections import Counter
from dataclasses import dataclass, field


def _w(s: str) -> frozenset[str]:
    return frozenset(s.split())


# Compact function-word lists (articles, pronouns, prepositions, auxiliaries) per language.
STOP_WORDS: dict[str, frozenset[str]] = {
    "en": _w("a an the and or but if then else when at by for with about against between into through during before "
             "after above below to from up down in out on off over under again further once here there why how all "
             "any both each few more most other some such no nor not only own same so than too very s t can will "
             "just don should now i me my myself we our ours ourselves you your yours yourself yourselves he him his "
             "himself she her hers herself it its itself they them their theirs themselves what which who whom this "
             "that these those am is are was were be been being have has had having do does did doing would could "
             "shall may might must of as while because until where also"),
    "ko": _w("이 그 저 것 수 등 들 및 에 에서 의 를 을 은 는 이가 가 와 과 도 로 으로 에게 한 하다 있다 되다 않다 없다 같다 "
             "그리고 그러나 그래서 또는 또한 하지만 때문에 위해 대한 통해 대해 매우 더 가장 좀 잘 못 안 왜 어떻게 무엇 누구 "
             "언제 어디 여기 거기 저기 우리 나 너 저 그녀 그들 이것 그것 저것 있는 하는 된 할 합니다 입니다 습니다"),
    "ja": _w("の に は を た が で て と し れ さ ある いる も する から な こと として い や れる など なっ ない この ため その "
             "あっ よう また もの という あり まで られ なる へ か だ これ によって により おり より による ず なり られる "
             "において ば なかっ なく しかし について せ だっ その後 できる それ う ので なお のみ でき き つ における および "
             "いう さらに でも ら たり その他 に関する たち ます ん なら に対して 特に せる 及び これら とき では にて ほか "
             "ながら うち そして とともに ただし かつて それぞれ または お ほど ものの に対する ほとんど と共に といった です"),
    "zh": _w("的 了 在 是 我 有 和 就 不 人 都 一 一个 上 也 很 到 说 要 去 你 会 着 没有 看 好 自己 这 那 他 她 它 们 我们 "
             "你们 他们 这个 那个 什么 怎么 为什么 哪 哪里 谁 吗 呢 吧 啊 与 及 或 但 但是 因为 所以 如果 虽然 而 而且 并 "
             "并且 被 把 对 从 向 以 于 为 之 其 此 这些 那些 可以 能 将 还 又 只 最 更 已 已经 等 等等 地 得"),
    "es": _w("de la que el en y a los del se las por un para con no una su al lo como más pero sus le ya o este sí "
             "porque esta entre cuando muy sin sobre también me hasta hay donde quien desde todo nos durante todos "
             "uno les ni contra otros ese eso ante ellos e esto mí antes algunos qué unos yo otro otras otra él tanto "
             "esa estos mucho quienes nada muchos cual poco ella estar estas algunas algo nosotros mi mis tú te ti tu "
             "tus ellas es son fue era ser está están ha han he"),
    "de": _w("der die das und in zu den von mit sich des auf für ist im dem nicht ein eine als auch es an werden aus "
             "er hat dass sie nach wird bei einer um am sind noch wie einem über einen so zum war haben nur oder aber "
             "vor zur bis mehr durch man sein wurde sei ich du wir ihr mir mich dir dich uns euch ihm ihn ihnen mein "
             "dein unser kein keine wenn dann weil was wer wo warum dieser diese dieses jener hier dort kann können "
             "muss müssen soll sollen hatte hatten wäre"),
    "fr": _w("de la le et les des en un du une que est pour qui dans a par plus pas au sur ne se ce il sont avec ou "
             "son lui nous comme mais on si leur y dont elle tout aux ces ses être cette fait été avoir sans sous "
             "je tu vous ils elles me te moi toi mon ma mes ton ta tes notre nos votre vos leurs quel quelle quand "
             "où pourquoi comment très aussi donc car ni ça cela ceci était ont avait suis es sommes êtes"),
    "pt": _w("de a o que e do da em um para é com não uma os no se na por mais as dos como mas foi ao ele das tem à "
             "seu sua ou ser quando muito há nos já está eu também só pelo pela até isso ela entre era depois sem "
             "mesmo aos ter seus quem nas me esse eles estão você tinha foram essa num nem suas meu às minha têm numa "
             "pelos elas havia seja qual será nós tenho lhe deles essas esses pelas este fosse dele tu te vocês vos"),
    "ru": _w("и в во не что он на я с со как а то все она так его но да ты к у же вы за бы по только ее мне было "
             "вот от меня еще нет о из ему теперь когда даже ну вдруг ли если уже или ни быть был него до вас нибудь "
             "опять уж вам ведь там потом себя ничего ей может они тут где есть надо ней для мы тебя их чем была сам "
             "чтоб без будто чего раз тоже себе под будет ж тогда кто этот того потому этого какой совсем ним здесь "
             "этом один почти мой тем чтобы нее были куда зачем всех можно при об это эта эти"),
    "ar": _w("في من على إلى عن مع هذا هذه ذلك تلك التي الذي الذين هو هي هم هن أنا نحن أنت أنتم كان كانت يكون تكون "
             "قد لقد لا لم لن ما ماذا متى أين كيف لماذا هل أو و ثم بل لكن إن أن إذا لو حتى كل بعض غير بين عند عندما "
             "بعد قبل حيث كما أي أيضا فقط جدا هناك هنا له لها لهم به بها فيه فيها منه منها عليه عليها إليه ب ل ك ف"),
    "hi": _w("का की के में है हैं को से पर और या यह वह ये वे इस उस एक ने भी तो ही था थी थे हो होता होती होते कर करना "
             "किया करते रहा रही रहे गया गई गए लिए साथ तक बाद पहले अब जब तब कब क्यों कैसे क्या कौन कहाँ मैं हम तुम आप "
             "मेरा हमारा तुम्हारा आपका उसका इसका उनका अपना कुछ कोई सब बहुत नहीं न मत जो जिस जिन कि अगर लेकिन क्योंकि "
             "इसलिए फिर वाला वाली वाले द्वारा बारे"),
    "th": _w("ที่ และ ใน ของ เป็น การ มี ได้ ว่า จะ ไม่ ให้ กับ นี้ ก็ แต่ หรือ โดย จาก ไป มา อยู่ ความ อย่าง ซึ่ง ด้วย "
             "นั้น เมื่อ ถึง แล้ว คือ ยัง ต้อง กัน ขึ้น ผู้ เพื่อ อีก ทั้ง เขา เรา ฉัน คุณ มัน พวก อะไร ทำไม อย่างไร "
             "ที่ไหน เมื่อไร ใคร มาก น้อย ทุก บาง ทั้งหมด เพราะ ถ้า แม้ ดังนั้น"),
    "tr": _w("ve bir bu da de için ile olarak çok daha ama en gibi ne o var mi mı mu mü ya hem ki kadar sonra önce "
             "her şey ben sen biz siz onlar benim senin onun bizim sizin onların bana sana ona bize size onlara beni "
             "seni onu bizi sizi onları şu şunlar bunlar ise veya ya da çünkü eğer değil yok olan oldu olur olmak "
             "etmek yapmak neden nasıl nerede kim hangi tüm bazı hiç"),
    "vi": _w("và của là có trong cho không được với các một những này đó khi đã sẽ đang từ đến như về tại bởi vì "
             "nên nhưng hoặc hay nếu thì mà cũng rất hơn nhất lại ra vào lên xuống tôi bạn anh chị em chúng ta họ "
             "nó ai gì đâu nào sao thế vậy đây kia ấy rồi còn chỉ mỗi mọi tất cả nhiều ít theo trên dưới sau trước"),
    "id": _w("yang dan di ke dari untuk pada dengan ini itu adalah tidak akan atau juga oleh sebagai dalam ada saya "
             "kami kita anda dia mereka ia nya telah sudah belum sedang bisa dapat harus karena jika maka tetapi tapi "
             "namun serta bahwa agar supaya seperti lebih sangat paling hanya saja lagi pun apa siapa kapan dimana "
             "mengapa bagaimana semua setiap beberapa banyak sedikit para antara atas bawah setelah sebelum saat"),
}


def get_stop_words(language: str | None = None) -> frozenset[str]:
    return STOP_WORDS.get(language or "en", STOP_WORDS["en"])


def remove_stop_words(tokens: list[str], language: str | None = None) -> list[str]:
    sw = get_stop_words(language)
    return [t for t in tokens if t.lower() not in sw]


# ------------------------------------------------------------------ synonym expansion
_SYNONYMS: dict[str, list[str]] = {
    "error": ["exception", "bug", "issue", "fault"], "bug": ["error", "defect", "issue"],
    "api": ["endpoint", "interface", "service"], "database": ["db", "datastore", "storage"],
    "db": ["database", "datastore"], "function": ["method", "procedure", "routine"],
    "method": ["function", "procedure"], "server": ["backend", "service", "host"],
    "client": ["frontend", "consumer", "user"], "config": ["configuration", "settings", "options"],
    "configuration": ["config", "settings"], "test": ["spec", "unittest", "testing"],
    "deploy": ["release", "ship", "publish"], "install": ["setup", "configure"],
    "search": ["query", "find", "lookup"], "async": ["asynchronous", "concurrent"],
    "sync": ["synchronous", "blocking"], "cache": ["buffer", "memoize"],
    "auth": ["authentication", "authorization", "login"], "docs": ["documentation", "manual", "guide"],
    "performance": ["speed", "latency", "throughput"],
}


def expand_query(query: str, *, max_expansions: int = 3) -> list[str]:
    """Extra terms (synonyms of the query's words), at most ``2 * max_expansions``."""
    words = query.lower().split()
    extra: list[str] = []
    for w in words:
        for syn in _SYNONYMS.get(w, ())[:max_expansions]:
            if syn not in words and syn not in extra:
                extra.append(syn)
    return extra[:max_expansions * 2]


# ------------------------------------------------------------------ typo correction
def _edit_distance(a: str, b: str) -> int:
    if len(a) < len(b):
        a, b = b, a
    if not b:
        return len(a)
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(cur[j - 1] + 1, prev[j] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


edit_distance = _edit_distance


def did_you_mean(query: str, vocabulary: list[str], *, max_distance: int = 2, max_suggestions: int = 3) -> list[str]:
    vocab = set(vocabulary)
    out: list[str] = []
    for tok in query.lower().split():
        if tok in vocab:
            continue
        best: tuple[int, str] | None = None
        for word in vocabulary:
            if abs(len(word) - len(tok)) > max_distance:
                continue
            d = _edit_distance(tok, word)
            if 0 < d <= max_distance and (best is None or (d, word) < best):
                best = (d, word)
        if best:
            fixed = query.replace(tok, best[1])
            if fixed != query and fixed not in out:
                out.append(fixed)
    return out[:max_suggestions]


# ------------------------------------------------------------------ natural-language filters
_DATE_RULES: tuple[tuple[re.Pattern[str], int | None], ...] = (
    (re.compile(r"\brecent(?:ly)?\b", re.I), 7),
    (re.compile(r"\blast\s+(\d+)\s+days?\b", re.I), None),
    (re.compile(r"\blast\s+week\b", re.I), 7), (re.compile(r"\blast\s+month\b", re.I), 30),
    (re.compile(r"\blast\s+year\b", re.I), 365), (re.compile(r"\btoday\b", re.I), 1),
    (re.compile(r"\byesterday\b", re.I), 2), (re.compile(r"\bthis\s+week\b", re.I), 7),
    (re.compile(r"\bthis\s+month\b", re.I), 30), (re.compile(r"\bthis\s+year\b", re.I), 365),
)
_DOMAIN = re.compile(r"\bsite:(\S+)\b|\bfrom\s+([\w.-]+\.(?:com|org|net|io|dev|edu|gov))\b", re.I)
_LANG_MAP = {"english": "en", "korean": "ko", "japanese": "ja", "chinese": "zh", "spanish": "es", "german": "de",
             "french": "fr", "portuguese": "pt", "russian": "ru", "arabic": "ar", "hindi": "hi", "thai": "th",
             "turkish": "tr", "vietnamese": "vi", "indonesian": "id", "malay": "ms"}
_LANG = re.compile(r"\bin\s+(" + "|".join(_LANG_MAP) + r")\b", re.I)


@dataclass
class ParsedQuery:
    cleaned_query: str
    date_from: float | None = None
    date_to: float | None = None
    include_domains: list[str] = field(default_factory=list)
    exclude_domains: list[str] = field(default_factory=list)
    language: str | None = None
    original_query: str = ""


def parse_natural_query(query: str) -> ParsedQuery:
    """Pull date ranges ("last 3 days"), ``site:`` / "from x.com" and "in korean" hints out of the query."""
    res = ParsedQuery(cleaned_query=query, original_query=query)
    text = query
    now = time.time()
    for pat, days in _DATE_RULES:
        m = pat.search(text)
        if not m:
            continue
        n = days if days is not None else int(m.group(1))
        res.date_from = now - n * 86400
        text = text[:m.start()] + text[m.end():]
    for m in _DOMAIN.finditer(text):
        dom = m.group(1) or m.group(2)
        if dom:
            res.include_domains.append(dom)
    text = _DOMAIN.sub("", text)
    m = _LANG.search(text)
    if m:
        res.language = _LANG_MAP.get(m.group(1).lower())
        text = text[:m.start()] + text[m.end():]
    res.cleaned_query = re.sub(r"\s+", " ", text).strip()
    return res


# ------------------------------------------------------------------ related searches
class RelatedSearchTracker:
    """Co-occurrence counts of query terms -> "related searches"."""

    def __init__(self, max_pairs: int = 10000):
        self._pairs: Counter[tuple[str, str]] = Counter()
        self._max_pairs = max_pairs

    def record(self, query: str) -> None:
        toks = sorted(set(query.lower().split()))
        for i, a in enumerate(toks):
            for b in toks[i + 1:]:
                self._pairs[(a, b)] += 1
        if len(self._pairs) > self._max_pairs:
            self._pairs = Counter(dict(self._pairs.most_common(self._max_pairs // 2)))

    def related(self, query: str, *, limit: int = 5) -> list[str]:
        toks = set(query.lower().split())
        cand: Counter[str] = Counter()
        for (a, b), n in self._pairs.items():
            if a in toks and b not in toks:
                cand[b] += n
            elif b in toks and a not in toks:
                cand[a] += n
        return [t for t, _ in cand.most_common(limit)]
