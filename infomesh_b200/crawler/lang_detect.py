"""Language identification without external models: Unicode script shares first, then function-word frequency for
Latin / Cyrillic languages (reference infomesh/crawler/lang_detect.py:302-445)."""
from __future__ import annotations

import re
from dataclasses import dataclass


@dataclass(frozen=True)
class LanguageDetection:
    language: str          # ISO 639-1; "en" with confidence 0.0 when undetermined (the reference's contract)
    confidence: float      # 0..0.95
    script: str = "Latin"  # Latin | Hangul | Kana | CJK | Cyrillic | Arabic | Devanagari | Thai | Greek | Hebrew | Unknown


LanguageDetectionResult = LanguageDetection      # name used by the reference (crawler/lang_detect.py:15)


_SCRIPTS: tuple[tuple[str, str, re.Pattern[str]], ...] = (
    ("ko", "Hangul", re.compile(r"[가-힯ᄀ-ᇿ㄰-㆏]")),
    ("ja", "Kana", re.compile(r"[぀-ゟ゠-ヿ]")),
    ("zh", "CJK", re.compile(r"[一-鿿㐀-䶿]")),
    ("th", "Thai", re.compile(r"[฀-๿]")),
    ("ar", "Arabic", re.compile(r"[؀-ۿݐ-ݿ]")),
    ("hi", "Devanagari", re.compile(r"[ऀ-ॿ]")),
    ("ru", "Cyrillic", re.compile(r"[Ѐ-ӿ]")),
    ("el", "Greek", re.compile(r"[Ͱ-Ͽ]")),
    ("he", "Hebrew", re.compile(r"[֐-׿]")),
)
_WORD = re.compile(r"[^\W\d_]+", re.UNICODE)

_COMMON: dict[str, frozenset[str]] = {
    "en": frozenset("the of and to in is that it for was on are with as be this have from or by not but what all "
                    "were when we there can an your which their said if will each about how up out them then she".split()),
    "es": frozenset("de la que el en y a los del se las por un para con no una su al lo como más pero sus le ya o "
                    "este sí porque esta entre cuando muy sin sobre también me hasta hay donde quien desde".split()),
    "fr": frozenset("de la le et les des en un du une que est pour qui dans par plus pas au sur ne se ce il sont "
                    "avec son lui nous comme mais on ou si leur y dont elle tout aux ces être cette été".split()),
    "de": frozenset("der die und in den von zu das mit sich des auf für ist im dem nicht ein eine als auch es an "
                    "werden aus er hat dass sie nach wird bei einer um am sind noch wie einem über einen so zum".split()),
    "pt": frozenset("de a o que e do da em um para é com não uma os no se na por mais as dos como mas foi ao ele "
                    "das tem à seu sua ou ser quando muito há nos já está eu também só pelo pela até isso".split()),
    "it": frozenset("di e il la che in a per un è del non con le si una i da al sono come più ma lo ha dei nel "
                    "alla delle gli anche questo su se o della quando era ci essere tra tutti".split()),
    "nl": frozenset("de van het een en in is dat op te zijn voor met die niet aan er om ook als dan maar bij of "
                    "uit nog naar door over ze zich heeft worden tot deze kan wordt hij".split()),
    "tr": frozenset("ve bir bu da de için ile olarak çok daha ama en gibi ne o var mi mı ya hem ki kadar sonra "
                    "önce her şey ben sen biz değil yok olan oldu".split()),
    "vi": frozenset("và của là có trong cho không được với các một những này đó khi đã sẽ đang từ đến như về "
                    "tại bởi vì nên nhưng hoặc nếu thì mà cũng rất".split()),
    "id": frozenset("yang dan di ke dari untuk pada dengan ini itu adalah tidak akan atau juga oleh sebagai dalam "
                    "ada saya kami kita anda dia mereka telah sudah bisa dapat harus karena".split()),
    "pl": frozenset("w i z na nie się do że to jest o jak a po co ale od za przez dla ich tak był być może tylko "
                    "który która które tym tego już jeszcze czy gdy bardzo także oraz przy nad pod".split()),
    "sv": frozenset("och i att det som en på är av för med till den har de inte om ett han men var sig från vi så "
                    "kan när hon skulle också eller efter vid nu än över under mycket hade här".split()),
    "da": frozenset("og i at det en den til er som på de med han af for ikke der var mig sig men et har om vi min "
                    "havde ham hun nu over da fra du ud sin dem os op man hans hvor eller hvad skal selv".split()),
    "no": frozenset("og i det er på en som til for av at med har de ikke den om et var fra men han seg vi kan så "
                    "ble hun eller også etter ved nå skal over under mye hadde være blir bare noe".split()),
    "fi": frozenset("ja on ei se että oli hän mutta joka kun niin myös kuin tai ovat olla tämä sen mitä vain jos "
                    "hänen sitä vielä nyt sekä jo minä sinä me te he ole kanssa mukaan jälkeen".split()),
    "cs": frozenset("a v se na je že to s z do o i pro ale jako za po by od tak jsou jeho být nebo jsem když už "
                    "který která které jen má byl byla bylo při pod nad před také velmi".split()),
    "ro": frozenset("și de în la a că cu pe nu este un o se din pentru care mai al ce sau dar au fost prin după "
                    "lui ale sunt fi această acest foarte când între până fără către deja".split()),
    "hu": frozenset("a az és hogy nem is egy ez volt van de meg csak mint már még vagy ha ki mi el be fel le én te "
                    "ő mert nagyon után között alatt által szerint amely amikor lehet kell".split()),
    "ru": frozenset("и в не на я что он с как а то все она так его но да ты к у же вы за бы по только ее мне "
                    "было вот от меня еще нет о из ему".split()),
    "uk": frozenset("і в не на я що він з як а то все вона так його але та ти до у ж ви за би по тільки її мені "
                    "було ось від мене ще немає про із йому це є".split()),
}
_CYRILLIC = ("ru", "uk")


_UNDETERMINED = LanguageDetection("en", 0.0, "Unknown")
MAX_CONFIDENCE = 0.95


def detect_language(text: str, *, min_text_length: int = 20) -> LanguageDetection:
    """Texts shorter than ``min_text_length`` are not judged at all (reference crawler/lang_detect.py:346-362)."""
    if not text or len(text) < min_text_length or not text.strip():
        return _UNDETERMINED
    sample = text[:5000]
    letters = sum(1 for ch in sample if ch.isalpha())
    if letters == 0:
        return _UNDETERMINED
    shares = {lang: (script, len(rx.findall(sample)) / letters) for lang, script, rx in _SCRIPTS}
    # Japanese text mixes kana with han: any meaningful kana share wins over "zh"
    if shares["ja"][1] >= 0.05:
        return LanguageDetection("ja", min(MAX_CONFIDENCE, 0.6 + shares["ja"][1] + shares["zh"][1]), "Kana")
    best_lang, (best_script, best_share) = max(shares.items(), key=lambda kv: kv[1][1])
    if best_share >= 0.3 and best_lang != "ru":
        return LanguageDetection(best_lang, min(MAX_CONFIDENCE, 0.5 + best_share / 2), best_script)
    words = [w.lower() for w in _WORD.findall(sample)]
    if not words:
        return _UNDETERMINED
    cyrillic = best_lang == "ru" and best_share >= 0.3
    pool = _CYRILLIC if cyrillic else [k for k in _COMMON if k not in _CYRILLIC]
    hits = {lang: sum(1 for w in words if w in _COMMON[lang]) for lang in pool}
    lang = max(hits, key=lambda k: hits[k])
    total = hits[lang]
    if total == 0:
        return LanguageDetection("ru", 0.35, "Cyrillic") if cyrillic else LanguageDetection("en", 0.1, "Latin")
    ranked = sorted(hits.values(), reverse=True)
    margin = (ranked[0] - ranked[1]) / ranked[0] if len(ranked) > 1 else 1.0
    density = min(1.0, total / max(len(words), 1) * 4)
    return LanguageDetection(lang, round(min(MAX_CONFIDENCE, 0.25 + 0.45 * density + 0.3 * margin), 3),
                             "Cyrillic" if cyrillic else "Latin")
